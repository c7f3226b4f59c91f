import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle
    pyoracle.lib()          # builds liboracle.so on first use
    return pyoracle


@pytest.fixture(scope="session")
def small_case(golden):
    """Seeded synthetic reads + what the reference's own binaries produced from them (tests/golden/make_golden.py)."""
    from nanosnp_b200.reads import PackedReads
    z = np.load(golden / "s1_small.npz")
    reads = PackedReads(z["pos"], z["flag"], z["mapq"], z["cigar_off"], z["cigar"], z["seq_off"], z["seq2"], z["nmask"])
    return {"ref": z["ref"], "reads": reads, "site_pos": z["site_pos"], "site_refbase": z["site_refbase"],
            "windows": z["windows"].astype(np.int32), "rows": int(z["mpileup_rows"])}


@pytest.fixture(scope="session")
def golden_weights(golden):
    from oracle.s2_restate import load_weights_npz
    return load_weights_npz(golden / "ont_pileup_weights.npz")


def oracle_s1(orc, reads, ref, tmpdir, contig="ctg1", **kw):
    mp = os.path.join(str(tmpdir), contig + ".mpileup")
    orc.mpileup_text(reads, contig, mp)
    return orc.s1_restate(mp, contig, ref, **kw)
