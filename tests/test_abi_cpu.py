"""The C-ABI library loads, exports every symbol include/walker_b200.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest


def test_header_declares_the_documented_surface(wb):
    syms = wb.declared_symbols()
    for must in ["wb_env_create", "wb_env_step", "wb_env_step_objects", "wb_env_take_actions", "wb_env_observe", "wb_env_set_state",
                 "wb_env_get_state", "wb_env_debug_contacts", "wb_material_register", "wb_policy_create", "wb_policy_forward",
                 "wb_policy_sample", "wb_ppo_grad", "wb_adam_step", "wb_policy_grad_buffer", "wb_last_error", "wb_init"]:
        assert must in syms


def test_library_exports_every_declared_symbol(wb):
    L = C.CDLL(wb.lib()._name)
    missing = [s for s in wb.declared_symbols() if not hasattr(L, s)]
    assert not missing, f"libwalker_b200.so does not export {missing}"


def test_no_torch_types_in_signatures(wb):
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "walker_b200.h")).read()
    assert 'extern "C"' in text
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "Tensor" not in code


def test_header_cites_reference_lines(wb):
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "walker_b200.h")).read()
    assert len(re.findall(r"\.cs:\d+", text)) >= 30


def test_compute_fails_loudly_without_a_gpu(wb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(wb.WalkerB200Error) as ei:
        wb.init(0)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(wb.WalkerB200Error):
        wb.EnvBatch(8)
    with pytest.raises(wb.WalkerB200Error):
        wb.PPOAgent()


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ppo-bipedalwalker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cs")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                bad = [ln for ln in text.splitlines()
                       if re.search(r"^\s*(import|from)\s+(oracle|np_oracle)\b|walker_oracle\.h|libwalker_oracle", ln)]
                assert not bad, f"{f} references the oracle: {bad}"


def test_hyperparameter_defaults_match_the_reference(wb):
    hp = wb.Hyperparams()
    assert wb.lib().wb_hyperparams_default(C.byref(hp)) == 0
    assert (hp.iterations, hp.max_timesteps, hp.batch_size, hp.use_gae, hp.normalize_advantages) == (50, 1000, 64, 0, 0)
    f = lambda x: float(__import__("numpy").float32(x))
    assert (hp.alpha, hp.beta1, hp.beta2, hp.adam_epsilon) == (f(0.001), f(0.9), f(0.999), f(1e-8))
    assert (hp.gamma, hp.lambda_, hp.epsilon, hp.log_std) == (f(0.9), f(0.95), f(0.3), -1.0)


def test_material_registry_is_host_side(wb):
    out = C.c_int32(-1)
    assert wb.lib().wb_material_register(3.0, 0.4, 0.6, C.byref(out)) == 0 and out.value >= 8
    im, e, mu = C.c_float(), C.c_float(), C.c_float()
    assert wb.lib().wb_material_get(out.value, C.byref(im), C.byref(e), C.byref(mu)) == 0
    assert (im.value, round(e.value, 6), round(mu.value, 6)) == (3.0, 0.4, 0.6)
    assert wb.lib().wb_material_get(1, C.byref(im), C.byref(e), C.byref(mu)) == 0 and im.value == 20.0  # Wood
    assert wb.lib().wb_material_get(63, C.byref(im), C.byref(e), C.byref(mu)) != 0


def _build_c_example(tmp_path):
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "ppo-bipedalwalker_b200", "lib")
    exe = str(tmp_path / "c_abi_smoke")
    cc = shutil.which("gcc") or shutil.which("cc")
    cmd = [cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
           os.path.join(root, "examples", "c_abi_smoke.c"), "-o", exe, "-L", libdir, "-lwalker_b200", f"-Wl,-rpath,{libdir}"]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert res.returncode == 0, res.stderr
    return exe


def test_header_is_plain_c99_and_a_c_program_links_and_is_refused_without_a_gpu(wb, tmp_path):
    """include/walker_b200.h compiles as strict C99 (-Wall -Wextra -Werror -pedantic) and examples/c_abi_smoke.c links against the
    library with no CUDA or C++ on its side; without an sm_100 device the program is refused loudly (exit 3: WB_ERR_NO_DEVICE)."""
    import subprocess
    import torch
    exe = _build_c_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present (tests/test_physics_gpu.py runs the program)")
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 3 and "no CPU fallback" in res.stderr


def _csharp_dir():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ppo-bipedalwalker_b200", "csharp")


def test_csharp_binding_declares_exactly_the_header_exports(wb):
    """ppo-bipedalwalker_b200/csharp/WalkerB200Native.cs holds one [DllImport] per function of include/walker_b200.h (no more, no
    fewer, apart from the documented pinned-pointer overload), generated by scripts/gen_pinvoke.py: regenerating it from the
    header reproduces the committed file byte for byte, and every declaration has the header's arity."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_pinvoke", os.path.join(root, "scripts", "gen_pinvoke.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = open(os.path.join(_csharp_dir(), "WalkerB200Native.cs")).read()
    assert text == gen.render(), "WalkerB200Native.cs is stale: run `python scripts/gen_pinvoke.py`"
    declared = re.findall(r"\[DllImport\(Lib\)\] public static extern \w+ (wb_[a-z0-9_]+)\(([^)]*)\);", text)
    names = [n for n, _ in declared]
    assert sorted(names) == wb.declared_symbols() and len(set(names)) == len(names)
    arity = {name: len(params) for _, name, params in gen.prototypes()}
    for name, params in declared:
        assert len([p for p in params.split(",") if p.strip()]) == arity[name], name
    overloads = re.findall(r'EntryPoint = "(wb_[a-z0-9_]+)"', text)
    assert overloads == ["wb_env_step"]


def test_csharp_structs_mirror_the_c_layouts(wb):
    """The blittable C# structs list the same number of 4-byte fields, in the same order of types, as the C structs."""
    header = open(os.path.join(os.path.dirname(_csharp_dir()), "..", "include", "walker_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    cs = open(os.path.join(_csharp_dir(), "WalkerB200Native.cs")).read()

    def c_fields(name):
        body = re.search(r"typedef struct \{([^}]*)\} " + name + ";", header).group(1)
        out = []
        for ctype, names in re.findall(r"(int32_t|uint32_t|float)\s+([^;]+);", body):
            out += [{"int32_t": "int", "uint32_t": "uint", "float": "float"}[ctype]] * len(names.split(","))
        return out

    def cs_fields(name):
        body = re.search(r"public struct " + name + r"\s*(?://[^\n]*)?\s*\{(.*?)\n?\}", cs, flags=re.S).group(1)
        out = []
        for ctype, names in re.findall(r"public (int|uint|float) ([^;(]+);", body):
            out += [ctype] * len(names.split(","))
        return out

    for c_name, cs_name in [("wb_hyperparams", "WbHyperparams"), ("wb_body_desc", "WbBodyDesc"), ("wb_joint_desc", "WbJointDesc"),
                            ("wb_pair_trace", "WbPairTrace"), ("wb_joint_trace", "WbJointTrace")]:
        assert c_fields(c_name) == cs_fields(cs_name), (c_name, c_fields(c_name), cs_fields(cs_name))
    assert C.sizeof(wb.Hyperparams) == 4 * len(c_fields("wb_hyperparams"))


def test_csharp_shim_covers_the_reference_surface():
    """The shim classes keep the reference's method names for this path: Environment.Update(float) / StepObjects /
    InitialState / GetConsoleInformation, the Walker and PPOAgent surfaces, and the two plugin adapters; every wb_* call they make
    exists in the generated binding."""
    d = _csharp_dir()
    env = open(os.path.join(d, "GpuEnvironment.cs")).read()
    for sig in ["public void Update(float deltaTime)", "public void StepObjects(float deltaTime)", "public void InitialState()",
                "GetConsoleInformation()", "public GpuWalker Walker(int index)"]:
        assert sig in env, sig
    walker = open(os.path.join(d, "GpuWalker.cs")).read()
    for sig in ["GetActions(PPO.Matrix state, out PPO.Matrix logProbabilities)", "TakeActions(PPO.Matrix actions)",
                "Train(Trajectory trajectory, Renderer renderer)", "GetState()", "GetJoints()", "GetPosition()", "GetChangeInPosition()",
                "public bool Terminal", "Reset()", "CreateCreature()", "Update()"]:
        assert sig in walker, sig
    agent = open(os.path.join(d, "GpuPPOAgent.cs")).read()
    for sig in ["public GpuPPOAgent(int stateSize, int actionSize)", "SampleActions(Matrix state, out Matrix logProbabilities, out Matrix mean, out Matrix std)",
                "public void Train(Trajectory trajectory, Renderer renderer)", "public void Save()", "public void Load(string type)"]:
        assert sig in agent, sig
    adapters = open(os.path.join(d, "Adapters.cs")).read()
    assert "IdOf(IMaterial material)" in adapters and "wb_material_register" in adapters and "wb_scene_create" in adapters
    binding = open(os.path.join(d, "WalkerB200Native.cs")).read()
    bound = set(re.findall(r"extern \w+ (wb_[a-z0-9_]+)\(", binding))
    for text in (env, walker, agent, adapters):
        for call in re.findall(r"Wb\.(wb_[a-z0-9_]+)\(", text):
            assert call in bound, call
