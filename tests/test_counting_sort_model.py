"""CPU model of the accumulation's stable counting sort (csrc/som_accumulate.cu: csort_hist / csort_prefix /
csort_offsets / csort_scatter): position = offsets[key] + (patches of the key in earlier blocks) + (patches of the key
earlier in the block: the bin's count when the patch's warp takes its turn + the rank inside the warp's match group).
Pure numpy against a stable argsort -- the order inside a unit must stay ascending by patch, which is what keeps the
segmented sums of the update in a fixed order."""
import numpy as np
import pytest

THREADS = 1024


def _counting_sort(keys, num_units, per):
    n = keys.size
    blocks = (n + per - 1) // per
    hist = np.zeros((blocks, num_units), dtype=np.int64)
    for b in range(blocks):                                     # csort_hist
        np.add.at(hist[b], keys[b * per:(b + 1) * per], 1)
    total = hist.sum(axis=0)
    prefix = np.cumsum(hist, axis=0) - hist                     # csort_prefix: exclusive over the blocks, per key
    offsets = np.concatenate([[0], np.cumsum(total)])           # csort_offsets
    skey = np.full(n, -1, dtype=np.int64)
    sid = np.full(n, -1, dtype=np.int64)
    for b in range(blocks):                                     # csort_scatter
        bins = np.zeros(num_units, dtype=np.int64)
        for u in range(0, per, THREADS):                        # rounds
            for warp in range(THREADS // 32):                   # warps take turns in warp order
                lanes = [b * per + u + warp * 32 + lane for lane in range(32)]
                lanes = [p for p in lanes if p < n and p < (b + 1) * per]
                seen = {}
                for p in lanes:                                 # match groups: rank among the lower lanes of the same key
                    k = int(keys[p])
                    rank_in_warp = seen.get(k, 0)
                    pos = offsets[k] + prefix[b, k] + bins[k] + rank_in_warp
                    skey[pos], sid[pos] = k, p
                    seen[k] = rank_in_warp + 1
                for k, c in seen.items():                       # the group's last lane adds the group size
                    bins[k] += c
    return skey, sid, offsets


@pytest.mark.parametrize("n,num_units,per,pattern", [(5000, 300, 2048, "random"), (4097, 16, 2048, "random"),
                                                     (3000, 50, 1024, "one_unit"), (6200, 1000, 3072, "runs"),
                                                     (33, 7, 2048, "random")])
def test_model_is_a_stable_sort(n, num_units, per, pattern):
    rng = np.random.default_rng(n)
    if pattern == "one_unit":
        keys = np.full(n, 17, dtype=np.int64)
    elif pattern == "runs":
        keys = (np.arange(n) // 37 % num_units).astype(np.int64)
    else:
        keys = rng.integers(0, num_units, n)
    skey, sid, offsets = _counting_sort(keys, num_units, per)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(sid, order) and np.array_equal(skey, keys[order])
    assert np.array_equal(offsets[:-1], np.searchsorted(keys[order], np.arange(num_units), side="left"))
    assert offsets[-1] == n
