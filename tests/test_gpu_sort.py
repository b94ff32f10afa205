"""GPU: the grouping primitive of the training steps (rb2_sort_positions, csrc/bucket_sort.cuh) against a stable
torch sort.  Bit-exact: sorted keys and the positions (ties in ascending position)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(keys_u32, bits):
    from recbole_b200 import ops
    dev = torch.device("cuda:0")
    k = torch.from_numpy(keys_u32.astype(np.uint32).view(np.int32)).to(dev)
    ks, pos = ops.sort_positions(k, bits)
    torch.cuda.synchronize()
    want_k, want_p = torch.sort(torch.from_numpy(keys_u32.astype(np.int64)).to(dev), stable=True)
    got_k = ks.cpu().numpy().view(np.uint32).astype(np.int64)
    got_p = pos.cpu().numpy().view(np.uint32).astype(np.int64)
    np.testing.assert_array_equal(got_k, want_k.cpu().numpy())
    np.testing.assert_array_equal(got_p, want_p.cpu().numpy())


@pytest.mark.parametrize("n,bits", [(1, 5), (31, 3), (257, 12), (2048, 13), (5000, 24), (100003, 17),
                                    (1 << 20, 24), (1 << 21, 21), (3000017, 26), (1 << 18, 32), (1 << 20, 1)])
def test_sort_positions_uniform(n, bits):
    rng = np.random.default_rng(n + bits)
    hi = (1 << bits)
    _check(rng.integers(0, hi, n, dtype=np.int64).astype(np.uint32), bits)


def test_sort_positions_skewed_and_constant():
    rng = np.random.default_rng(3)
    n = 1 << 21
    # Zipf item ids (a few keys with > 100k occurrences), a table of 2M rows
    z = np.minimum(np.exp(rng.random(n) * np.log(2_000_000)).astype(np.int64), 1_999_999)
    perm = rng.permutation(2_000_001)
    _check(perm[z].astype(np.uint32), 21)
    _check(np.full(300_000, 12345, dtype=np.uint32), 20)          # one run
    _check(np.arange(400_000, dtype=np.uint32)[::-1].copy(), 19)    # reversed
    # a field with 3 values next to huge ones (FM): 3 passes, 26 bits
    k = np.where(rng.random(n) < 0.5, rng.integers(0, 3, n), rng.integers(0, 1 << 26, n)).astype(np.uint32)
    _check(k, 26)


def test_sort_positions_many_elements_wide_counters():
    """> 65535 elements per warp slice: the 32-bit-counter instantiation."""
    rng = np.random.default_rng(5)
    n = 160_000_000 // 1          # chunk / 8 >= 65535 at <= 296 blocks
    k = rng.integers(0, 1 << 22, n, dtype=np.int64).astype(np.uint32)
    _check(k, 22)
