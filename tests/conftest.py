import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class G:
        clip = np.load(os.path.join(GOLDEN, "clip10s.npz"))
        c1 = np.load(os.path.join(GOLDEN, "c1_part0.npz"))
        c2 = np.load(os.path.join(GOLDEN, "c2_gap.npz"))
        c3 = np.load(os.path.join(GOLDEN, "c3_mask.npz"))
        f4 = np.load(os.path.join(GOLDEN, "f4_siblings.npz"))
        sr = 44100

        @staticmethod
        def gap_input_i16():
            """demo_assets/part2/damaged_gap.wav rebuilt: the clean clip with [centre-sr, centre+sr) zeroed
            (generate_part2_data.py:36-43; verified equal to the shipped file by make_golden.py)."""
            x = G.clip["original_i16"].copy()
            gs, ge = G.c2["gap"]
            x[gs:ge] = 0
            return x

        @staticmethod
        def mask_input_i16():
            keep = np.unpackbits(G.c3["keep_bits"])[: len(G.clip["original_i16"])].astype(bool)
            return np.where(keep, G.clip["original_i16"], 0).astype(np.int16)

    return G
