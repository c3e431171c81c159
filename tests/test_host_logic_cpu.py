"""Host-side logic that needs no GPU: network DSL, Xavier init, sharding, bench plumbing."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parse_layers_matches_reference_dsl(wb):
    assert wb.ParseLayers(wb.DEFAULT_ACTOR) == [(0, 64), (2, 0), (0, 64), (2, 0), (0, 4), (3, 0)]
    assert wb.ParseLayers(wb.DEFAULT_CRITIC) == [(0, 64), (2, 0), (0, 1)]
    assert wb.ParseLayers("Input |8| (ReLU) |2| Output") == [(0, 8), (1, 0), (0, 2)]
    for bad in ["", "Input Output", "Input |64| (Sigmoid) |1| Output", "input |64| Output", "Input |64| Output "]:
        with pytest.raises(ValueError):
            wb.ParseLayers(bad)


def test_xavier_shapes_and_scale(wb):
    from ppo_bipedalwalker_b200.ppo import dense_shapes
    layers = wb.ParseLayers(wb.DEFAULT_ACTOR)
    assert dense_shapes(12, layers) == [(64, 12), (64, 64), (4, 64)]
    flat = wb.xavier_flat(12, layers, np.random.default_rng(0))
    assert flat.size == 5252 and flat.dtype == np.float32
    w2 = flat[832:832 + 4096]
    assert abs(w2.std() - np.sqrt(2 / 128)) < 0.01 and not flat[768:832].any()  # biases start at zero


def test_shard_range_partitions_exactly():
    from ppo_bipedalwalker_b200.dist import shard_range
    for n, w in [(65536, 8), (4096, 2), (10, 4), (7, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["walkers_per_gpu"] == 4096 and line["higher_is_better"] is True


def test_bench_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0  # no CPU fallback, no silent oracle substitution
