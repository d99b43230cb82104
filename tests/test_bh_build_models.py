"""CPU: the index logic of the CTA-local Barnes-Hut build (tools/bh_local_model.py mirrors bh_emit_local_kernel and the climb on
symbolic values) against a plain recursion over the sorted keys, and the 8-ary owner search against a linear scan."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import bh_local_model as M  # noqa: E402


def test_local_windows_and_climb_reproduce_the_tree():
    random.seed(1)
    for trial in range(150):
        n = random.choice([1, 2, 3, 5, 17, 63, 64, 65, 100, 257, 300, 700])
        m, h = random.choice([4, 8, 16]), random.choice([4, 8, 16, 32])
        mode = random.randrange(4)
        if mode == 0:
            keys = [random.getrandbits(64) for _ in range(n)]
        elif mode == 1:
            keys = [random.getrandbits(12) << 52 for _ in range(n)]                       # many coincident bodies
        elif mode == 2:
            keys = [(random.getrandbits(6) << 58) | random.getrandbits(8) for _ in range(n)]   # deep single-child chains
        else:
            base = random.getrandbits(64)
            keys = [base ^ random.getrandbits(random.choice([4, 10, 30, 64])) for _ in range(n)]
        M.check(keys, m, h)


def test_owner_search_equals_linear_scan():
    random.seed(2)
    for trial in range(5000):
        n = random.choice([1, 2, 5, 9, 70, 600, 5000])
        bits = random.choice([3, 6, 10, 16])
        keys = sorted(random.getrandbits(bits) for _ in range(n))
        s = random.randrange(n)
        shp = random.randrange(0, bits)
        pp = keys[s] >> shp
        want = min(j for j in range(s + 1) if (keys[j] >> shp) >= pp)
        assert M.first_with_prefix(keys, s, shp, pp) == want
