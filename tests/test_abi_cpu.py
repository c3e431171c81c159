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
