"""Golden outputs of the reference's own post-processing scripts (only possible in the build container):

    python tests/golden/make_golden_post.py

Runs /root/reference/scripts/merge.py and select_high_quality_hetesnps.py (unmodified, as subprocesses) on
tests/golden/s2_small.vcf plus a seeded haplotype-model result table and stores inputs and outputs under tests/golden/post/.
"""
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/scripts")
OUT = HERE / "post"


def main():
    OUT.mkdir(exist_ok=True)
    rng = np.random.default_rng(2024)
    labels = ["AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT", "DD", "AD", "CD", "GD", "TD", "II", "AI", "CI", "GI", "TI", "ID"]
    vcf_lines = (HERE / "s2_small.vcf").read_text().splitlines(keepends=True)
    # a second contig and phased-style genotypes so that the per-contig split and the '|' handling are exercised
    extra = []
    for l in vcf_lines:
        if l.startswith("#"):
            continue
        f = l.rstrip("\n").split("\t")
        if rng.random() < 0.08:
            f[0] = "ctg2"
            s = f[9].split(":"); s[0] = s[0].replace("/", "|"); f[9] = ":".join(s)
            extra.append("\t".join(f) + "\n")
    vcf = OUT / "pileup.vcf"
    vcf.write_text("".join(vcf_lines) + "".join(extra))
    rows = []
    for l in vcf_lines + extra:
        if l.startswith("#"):
            continue
        f = l.split("\t")
        if float(f[5]) <= 19 and rng.random() < 0.7:
            rows.append(f"{f[0]}\t{f[1]}\t{labels[int(rng.integers(0, 21))]}\t{round(float(rng.uniform(0, 40)), 2)}\n")
    for k in range(40):                                   # haplotype calls at positions the pileup VCF does not have
        rows.append(f"ctg1\t{900000 + k}\tAC\t30.0\n")
    (OUT / "haplotype.csv").write_text("".join(rows))
    for q in (19, 15):
        subprocess.run([sys.executable, str(REF / "merge.py"), "--pileup_vcf", str(vcf), "--cat_predict", str(OUT / "haplotype.csv"),
                        "--quality", str(q), "--output", str(OUT / f"merge_q{q}.vcf")], check=True)
    (OUT / "empty.csv").write_text("")
    subprocess.run([sys.executable, str(REF / "merge.py"), "--pileup_vcf", str(vcf), "--cat_predict", str(OUT / "empty.csv"),
                    "--output", str(OUT / "merge_empty.vcf")], check=True, stdout=subprocess.DEVNULL)
    for q in (14, 16):
        d = OUT / f"split_q{q}"
        subprocess.run([sys.executable, str(REF / "select_high_quality_hetesnps.py"), "--pileup_vcf", str(vcf), "--support_quality", str(q),
                        "--output_dir", str(d)], check=True)
    print({p.name: p.stat().st_size for p in sorted(OUT.rglob("*")) if p.is_file()})


if __name__ == "__main__":
    main()
