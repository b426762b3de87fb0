"""CPU checks of bench.py's workload generator: the global synthetic ensemble is defined blockwise from the global conformation
index, so the shards the ranks build on their own concatenate to exactly the set an unsharded run builds (the basis of the
``multi_gpu_parity`` check of the strong-scaling bench), and the CPU arm sees the same conformations as the GPU arm."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def test_global_slice_shards_concatenate_to_the_whole():
    import bench
    from paramol_b200.Utils import dist as pdist
    rng = np.random.default_rng(0)
    xb = rng.normal(0, 1, (10, 14, 3))
    n = 3 * bench.BLOCK + 1234          # ragged last block
    whole = bench.global_slice(xb, 0, n, 14)
    for world in (2, 3, 8):
        parts = [bench.global_slice(xb, *pdist.shard_bounds(n, world, r), 14) for r in range(world)]
        for k in range(3):
            assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k]), (world, k)
    # a slice that starts and ends inside blocks
    lo, hi = bench.BLOCK + 17, 2 * bench.BLOCK + 5
    part = bench.global_slice(xb, lo, hi, 14)
    for k in range(3):
        assert np.array_equal(part[k], whole[k][lo:hi])


def test_flop_count_matches_survey_formula():
    from bench_configs import flops_per_conformation
    # aniline: 14 atoms, 14 bonds, 21 angles, 35 torsions, 56 interacting pairs -> 10 224 flop
    assert flops_per_conformation(14, 14, 21, 35, 56, term_type="components") == 10224
