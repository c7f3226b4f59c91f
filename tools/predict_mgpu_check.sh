# usage: bash tools/predict_mgpu_check.sh [N=2]   (needs N GPUs)
set -e
N=${1:-2}; D=/tmp/nsnp_mgpu
python tools/predict_mgpu_check.py prepare $D
ARGS="-config nanosnp_b200/config/ont_pileup.yaml -model_path tests/golden/ont_pileup_weights.npz -data $D/data -reference $D/ref.fa --region_len 100000"
python -m nanosnp_b200.predict $ARGS -output $D/one.vcf
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 -m nanosnp_b200.predict $ARGS -output $D/multi.vcf
python tools/predict_mgpu_check.py compare $D/one.vcf $D/multi.vcf
