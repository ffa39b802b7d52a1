"""K-ORDER, large bucket tables (fxg_collapse.cu: k_bucket_pos / k_touch_min_part / k_touch_key_part): the nodes of an epoch are
partitioned by the top bits of their bucket number before the first-touch pass.  Model of the claim the kernel relies on: a
STABLE sort by first-touch position gives the same node order whether it runs on the original sequence or on the stably
partitioned one, because equal keys share a bucket, hence a partition.  (The epoch rule itself — libstdc++ unordered_map,
fastx_collapser.cpp:116-122 via SURVEY Appendix B — is pinned by tests/test_collapser_order.py against the reference binary.)"""
import numpy as np
import pytest


def epoch_direct(nodes, bucket, B):
    m = len(nodes)
    touch = np.full(B, np.iinfo(np.int64).max, np.int64)
    np.minimum.at(touch, bucket, np.arange(m))
    key = touch[bucket]
    return nodes[np.argsort(key, kind="stable")]


def epoch_partitioned(nodes, bucket, B, part_bits):
    m = len(nodes)
    hb = max(int(B - 1).bit_length(), 1)
    lb = max(hb - part_bits, 0)
    part = np.argsort(bucket >> lb, kind="stable")            # one stable radix pass on the top bits
    bucket_p, pos_p, nodes_p = bucket[part], np.arange(m)[part], nodes[part]
    touch = np.full(B, np.iinfo(np.int64).max, np.int64)
    np.minimum.at(touch, bucket_p, pos_p)                     # atomicMin with the ORIGINAL positions
    key_p = touch[bucket_p]
    return nodes_p[np.argsort(key_p, kind="stable")]


@pytest.mark.parametrize("m,B,bits", [(1, 13, 8), (50, 13, 2), (1000, 1109, 3), (20000, 20753, 8), (20753, 20753, 8), (5000, 85229, 5)])
def test_partitioned_epoch_equals_direct(m, B, bits):
    rng = np.random.default_rng(m * 31 + B)
    nodes = rng.permutation(m).astype(np.int64)
    bucket = rng.integers(0, B, m)
    if m > 10:
        bucket[rng.integers(0, m, m // 3)] = bucket[rng.integers(0, m, m // 3)]      # more shared buckets
    assert np.array_equal(epoch_direct(nodes, bucket, B), epoch_partitioned(nodes, bucket, B, bits))
