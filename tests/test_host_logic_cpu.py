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
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the arm must still use all host threads (it passes the count explicitly)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["walkers_per_gpu"] == 4096 and line["higher_is_better"] is True


def test_bench_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0  # no CPU fallback, no silent oracle substitution


def test_hyperparameter_json_roundtrip_and_validation(wb, tmp_path):
    """Hyperparameters.SerializeJson / DeserializeJson / ValidateHyperparameterValues (Hyperparameters.cs:124-217)."""
    import json
    s = wb.Settings(Iterations=25, BatchSize=128, Gamma=0.97, UseGAE=True, FilePath="/data")
    path = tmp_path / "Data" / "settings.json"
    wb.SerializeJson(s, str(path))
    doc = json.loads(path.read_text())
    assert list(doc)[:6] == ["GameSpeed", "CollectData", "SaveWeights", "Iterations", "MaxTimesteps", "RoughFloor"]  # the reference's member order
    assert doc["CriticNeuralNetwork"] == "Input |64| (LeakyReLU) |1| Output" and doc["LogStandardDeviation"] == -1.0
    back = wb.DeserializeJson(str(path))
    assert back == s
    hp = back.to_hyperparams()
    assert (hp.iterations, hp.batch_size, hp.use_gae) == (25, 128, 1) and abs(hp.gamma - 0.97) < 1e-7
    # out-of-range value: logged, current settings kept (the reference logs and returns, :141-185)
    logged = []
    doc["Iterations"] = 200
    path.write_text(json.dumps(doc))
    kept = wb.DeserializeJson(str(path), current=s, log=logged.append)
    assert kept is s and "Invalid iterations count" in logged[0]
    # the boundary semantics are the reference's: Beta1 == 1 and Gamma == 1 are accepted, Epochs == 50 is not
    doc.update(Iterations=199, Beta1=1.0, Gamma=1.0)
    path.write_text(json.dumps(doc))
    assert wb.DeserializeJson(str(path)).Beta1 == 1.0
    doc["Epochs"] = 50
    path.write_text(json.dumps(doc))
    assert wb.DeserializeJson(str(path), current=s) is s
    # malformed document / missing members (System.Text.Json leaves 0 / false / null, which then fails validation)
    path.write_text("{ not json")
    assert wb.DeserializeJson(str(path), current=s, log=logged.append) is s and "JSON deserializer error" in logged[-1]
    path.write_text(json.dumps({"Iterations": 50}))
    assert wb.DeserializeJson(str(path), current=s) is s


def test_iobject_shape_factories_match_the_numpy_restatement(wb, O):
    """Square/Triangle/Hexagon/Pole.FromSize and Skeleton.SmoothCorners (Objects/RigidBodies/*.cs, Skeleton.cs:33-53): the host mirror
    that feeds wb_scene_create and the NumPy restatement produce the same float32 vertex lists."""
    import np_oracle as P
    m = (15, 0.3, 1.0)
    c = (123.4, 567.8)
    pairs = [(wb.Square.FromSize("Metal", c, 61.7), P.square_from_size(m, P.Vec(*c), 61.7)),
             (wb.Triangle.FromSize("Metal", c, 33.3), P.triangle_from_size(m, P.Vec(*c), 33.3)),
             (wb.Hexagon.FromSize("Metal", c, 80.1), P.hexagon_from_size(m, P.Vec(*c), 80.1)),
             (wb.Pole.FromSize("Metal", c, 75), P.pole_from_size(m, P.Vec(*c), 75, "p")),
             (wb.Hexagon.FromSize("Metal", c, 80.1).SmoothCorners(1), P.smooth_corners(P.hexagon_from_size(m, P.Vec(*c), 80.1), 1)),
             (wb.Square.FromSize("Metal", c, 10).SmoothCorners(2), P.smooth_corners(P.square_from_size(m, P.Vec(*c), 10), 2))]
    for o, b in pairs:
        ref = np.array([(v.x, v.y) for v in b.skeleton.vectors], np.float32)
        assert o.vertices.shape == ref.shape and np.array_equal(o.vertices.view(np.uint32), ref.view(np.uint32))
    assert pairs[-1][0].vertices.shape == (16, 2)  # the largest polygon a scene accepts


def test_reference_scene_pieces_match_the_numpy_restatement(wb, O):
    """CreateCreature (Walker.cs:40-46,155-209), CreateFloor and CreateRoughFloor (Environment.cs:211-261) as IObject lists: same
    vertices, list order, joint topology and association lists as the NumPy restatement / the reference source."""
    import np_oracle as P
    env = P.Environment()
    objs, joints = wb.CreateCreature()
    assert len(objs) == 5 and len(joints) == 4
    for o, b in zip(objs, env.dyn()):
        ref = np.array([(v.x, v.y) for v in b.skeleton.vectors], np.float32)
        assert np.array_equal(o.vertices.view(np.uint32), ref.view(np.uint32))
        assert sorted(objs.index(a) for a in o.associated) == sorted(env.dyn().index(a) for a in b.associated)
        assert o.acceleration == (0.0, 980.0)
    assert objs[2].inverseInertia == 0.0003 and all(o.inverseInertia < 0 for i, o in enumerate(objs) if i != 2)
    for j, rj in zip(joints, env.joints):
        assert (objs.index(j.bodyA), objs.index(j.bodyB), j.indexA, j.indexB) == (env.dyn().index(rj.a), env.dyn().index(rj.b), rj.ia, rj.ib)
    floor = wb.CreateFloor()
    assert len(floor) == 1 and floor[0].isStatic and floor[0].isFloor
    assert np.array_equal(floor[0].vertices, np.array([(v.x, v.y) for v in env.floor.skeleton.vectors], np.float32))
    # rough floor: Environment.cs:236-259 by hand for two segments
    heights = [7, 93, 0, 55, 12, 99, 31, 64, 8, 77, 40]
    rough = wb.CreateRoughFloor(heights)
    assert len(rough) == 10 and all(o.isStatic and o.isFloor and o.vertices.shape == (4, 2) for o in rough)
    assert rough[0].vertices.tolist() == [[-50, 1050], [-50, 807], [-50, 893], [70, 1050]]
    assert rough[1].vertices.tolist() == [[70, 1050], [-50, 893], [70, 800], [190, 1050]]
    assert rough[9].vertices.tolist() == [[1030, 1050], [910, 877], [1030, 840], [1150, 1050]]
    for bad in (lambda: wb.CreateRoughFloor(heights, segments=0), lambda: wb.CreateRoughFloor(heights[:-1]),
                lambda: wb.CreateRoughFloor([100] + heights[1:])):
        try:
            bad()
            raise AssertionError("accepted an invalid rough floor")
        except ValueError:
            pass


def test_refharness_glue_exports_and_compares_the_golden_rollout(tmp_path):
    """scripts/refharness_io.py: the input file for the C# RefHarness is well formed, and the comparer accepts exactly the golden
    states (and names the first differing field otherwise).  The harness itself needs a .NET SDK and is not run here."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "scripts", "refharness_io.py")
    inp = tmp_path / "golden_actions.txt"
    subprocess.run([sys.executable, script, "export", str(inp)], check=True, capture_output=True)
    lines = inp.read_text().splitlines()
    assert lines[0] == "8 90 50" and lines[1].split() == ["Ice", "Wood", "Paper", "Titanium", "Carpet", "Rubber", "Metal", "SuperRubber"]
    assert len(lines) == 2 + 90 * 8 and all(len(ln.split()) == 5 for ln in lines[2:])
    g = np.load(os.path.join(root, "tests", "golden", "physics_rollout.npz"))
    words = np.ascontiguousarray(g["states"], np.float32).view(np.uint32)

    def dump(a, path):
        with open(path, "w") as fh:
            for t in range(a.shape[0]):
                for e in range(a.shape[1]):
                    fh.write(" ".join(f"{x:08x}" for x in a[t, e]) + "\n")

    dump(words, tmp_path / "ok.txt")
    bad = words.copy()
    bad[10, 3, 70] ^= 1
    dump(bad, tmp_path / "bad.txt")
    ok = subprocess.run([sys.executable, script, "compare", str(tmp_path / "ok.txt")], capture_output=True, text=True)
    assert ok.returncode == 0 and "pinned" in ok.stdout
    res = subprocess.run([sys.executable, script, "compare", str(tmp_path / "bad.txt")], capture_output=True, text=True)
    assert res.returncode == 1 and "env-step 10, walker 3" in res.stdout and "LLU.velocity.x" in res.stdout


def test_weight_values_print_like_dotnet_float_tostring(wb):
    """Matrix.Save joins float.ToString() (Matrix.cs:133-136): shortest round-trip digits, decimal notation from 1e-4 up to 1e7,
    otherwise E+XX / E-XX -- so a .weights file written here is byte-compatible with one the reference writes."""
    import numpy as np
    from ppo_bipedalwalker_b200.ppo import net_float_str
    cases = {0.1: "0.1", -0.25: "-0.25", 1.0: "1", 0.0: "0", 1e-5: "1E-05", 9.999999e-6: "9.999999E-06", 1234567.0: "1234567",
             1e7: "1E+07", 12345678.0: "1.2345678E+07", 3.4028235e38: "3.4028235E+38", 0.10000000149011612: "0.1", 1.5e-45: "1E-45"}
    for value, text in cases.items():
        assert net_float_str(np.float32(value)) == text, (value, net_float_str(np.float32(value)), text)
    rng = np.random.default_rng(0)
    for x in (rng.normal(size=2000) * 10.0 ** rng.integers(-8, 9, 2000)).astype(np.float32):
        assert np.float32(float(net_float_str(x).replace("E", "e"))) == x  # round-trips exactly
