"""ctypes binding of the CPU ORACLE (oracle/_build/libwalker_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (ppo-bipedalwalker_b200/) never imports this module.
PARITY UNPINNED (see oracle/walker_oracle.h): no reference golden vectors exist and no C# toolchain is
available; the oracle is pinned by hand-derived known answers and an independent NumPy restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libwalker_oracle.so")

STATE_FLOATS = 92
STATE_INTS = 2
OBS = 12
ACT = 4
PAIR_SLOTS = 9
FLAG_TERMINAL = 1 << 5
FLAG_FLOOR_FIRST = 1 << 6

MATERIALS = ["Ice", "Wood", "Paper", "Titanium", "Carpet", "Rubber", "Metal", "SuperRubber"]
DT_FRAME = float(np.float32(0.0166667))  # (float)TimeSpan.FromTicks(166667).TotalSeconds, SURVEY Appendix C

PAIR_TRACE_DTYPE = np.dtype(
    [("other", "<i4"), ("aabb", "<i4"), ("sat", "<i4"), ("axis", "<i4"), ("nx", "<f4"), ("ny", "<f4"), ("depth", "<f4"),
     ("ncontacts", "<i4"), ("c0x", "<f4"), ("c0y", "<f4"), ("c1x", "<f4"), ("c1y", "<f4")]
)
JOINT_TRACE_DTYPE = np.dtype([("active", "<i4"), ("depth", "<f4")])


class Material(C.Structure):
    _fields_ = [("inverse_mass", C.c_float), ("restitution", C.c_float), ("friction", C.c_float)]


class Hyper(C.Structure):
    _fields_ = [("alpha", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("adam_epsilon", C.c_float),
                ("epsilon", C.c_float), ("log_std", C.c_float), ("gamma", C.c_float), ("lambda_", C.c_float),
                ("batch_size", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, strict fp32). Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in ("walker_oracle_physics.c", "walker_oracle_ppo.c", "walker_oracle.h", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int32)
        u8p = C.POINTER(C.c_uint8)
        vp = C.c_void_p
        L.wo_builtin_material.restype = C.POINTER(Material)
        L.wo_builtin_material.argtypes = [C.c_int]
        L.wo_env_sizeof.restype = C.c_int
        L.wo_env_init.argtypes = [vp, Material, Material]
        L.wo_env_reset.argtypes = [vp]
        L.wo_env_take_actions.argtypes = [vp, fp]
        L.wo_env_step_objects.argtypes = [vp, C.c_float, C.c_int, vp, vp]
        L.wo_env_observe.argtypes = [vp, C.c_int, fp, fp, u8p]
        L.wo_env_get_obs.argtypes = [vp, fp]
        L.wo_env_step.argtypes = [vp, fp, C.c_float, C.c_int, C.c_int, C.c_int, fp, fp, u8p]
        L.wo_env_get_state.argtypes = [vp, fp, ip]
        L.wo_env_set_state.argtypes = [vp, fp, ip]
        L.wo_batch_step.argtypes = [vp, C.c_int, fp, C.c_float, C.c_int, C.c_int, C.c_int, fp, fp, u8p, C.c_int]
        L.wo_sat.restype = C.c_int
        L.wo_sat.argtypes = [fp, C.c_int, fp, C.c_int, fp, fp, fp, fp, ip]
        L.wo_contacts.restype = C.c_int
        L.wo_contacts.argtypes = [fp, C.c_int, fp, C.c_int, fp, fp]
        L.wo_pole_from_size.argtypes = [C.c_float, C.c_float, C.c_float, fp, fp]
        L.wo_rotz.argtypes = [C.c_float, fp, fp]
        L.wo_hyper_defaults.argtypes = [C.POINTER(Hyper)]
        L.wo_net_create.restype = vp
        L.wo_net_create.argtypes = [C.c_int, ip, ip, C.c_int]
        L.wo_net_destroy.argtypes = [vp]
        L.wo_net_num_params.restype = C.c_int
        L.wo_net_num_params.argtypes = [vp]
        L.wo_net_output_size.restype = C.c_int
        L.wo_net_output_size.argtypes = [vp]
        for name in ("wo_net_set_params", "wo_net_get_params", "wo_net_get_grads"):
            getattr(L, name).argtypes = [vp, fp]
        L.wo_net_get_adam.argtypes = [vp, fp, fp, ip]
        L.wo_net_set_adam.argtypes = [vp, fp, fp, ip]
        L.wo_net_forward.argtypes = [vp, fp, fp, C.c_int]
        L.wo_net_feedback.argtypes = [vp, fp]
        L.wo_net_zero.argtypes = [vp]
        L.wo_net_optimise.argtypes = [vp, C.POINTER(Hyper)]
        L.wo_log_prob.restype = C.c_float
        L.wo_log_prob.argtypes = [C.c_float] * 3
        L.wo_box_muller.restype = C.c_float
        L.wo_box_muller.argtypes = [C.c_float] * 4
        L.wo_sample_actions.argtypes = [vp, C.POINTER(Hyper), fp, fp, fp, fp, fp]
        L.wo_ppo_train_batch.restype = C.c_int
        L.wo_ppo_train_batch.argtypes = [vp, vp, C.POINTER(Hyper), C.c_int, fp, fp, fp, fp, fp, C.c_int, fp, fp]
        L.wo_ppo_sample_grad.restype = C.c_int
        L.wo_ppo_sample_grad.argtypes = [C.POINTER(Hyper), C.c_int, fp, fp, fp, C.c_float, C.c_float, C.c_float, fp, fp]
        L.wo_mc_returns.argtypes = [fp, fp, C.c_int, C.c_float, fp, fp]
        L.wo_gae.argtypes = [fp, fp, C.c_int, C.c_float, C.c_float, fp, fp]
        L.wo_normalize.argtypes = [fp, C.c_int, C.c_float]
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _u8p(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def material(m) -> Material:
    """m: builtin id, builtin name, or (inverse_mass, restitution, friction)."""
    if isinstance(m, Material):
        return m
    if isinstance(m, str):
        m = MATERIALS.index(m)
    if isinstance(m, (int, np.integer)):
        return lib().wo_builtin_material(int(m)).contents
    inv_mass, restitution, friction = m
    return Material(float(inv_mass), float(restitution), float(friction))


class EnvBatch:
    """N independent reference environments (AoS array of wo_env)."""

    def __init__(self, n: int, floor="Metal", walker="Carpet"):
        L = lib()
        self.n = n
        self.size = L.wo_env_sizeof()
        self.buf = (C.c_char * (self.size * n))()
        self.base = C.addressof(self.buf)
        floors = floor if isinstance(floor, (list, tuple, np.ndarray)) and not _is_triple(floor) else [floor] * n
        walkers = walker if isinstance(walker, (list, tuple, np.ndarray)) and not _is_triple(walker) else [walker] * n
        for i in range(n):
            L.wo_env_init(self.ptr(i), material(floors[i]), material(walkers[i]))

    def ptr(self, i: int) -> int:
        return self.base + i * self.size

    def reset(self, i: int):
        lib().wo_env_reset(self.ptr(i))

    def take_actions(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, ACT)
        for i in range(self.n):
            lib().wo_env_take_actions(self.ptr(i), _fp(a[i]))

    def step_objects(self, dt: float, iterations: int, trace: bool = False):
        L = lib()
        pt = jt = None
        if trace:
            pt = np.zeros((self.n, iterations, PAIR_SLOTS), dtype=PAIR_TRACE_DTYPE)
            jt = np.zeros((self.n, iterations, 4), dtype=JOINT_TRACE_DTYPE)
        for i in range(self.n):
            L.wo_env_step_objects(self.ptr(i), dt, iterations, pt[i].ctypes.data if trace else None,
                                  jt[i].ctypes.data if trace else None)
        return pt, jt

    def observe(self, max_timesteps: int = 1000):
        obs = np.zeros((self.n, OBS), np.float32)
        rew = np.zeros(self.n, np.float32)
        done = np.zeros(self.n, np.uint8)
        for i in range(self.n):
            lib().wo_env_observe(self.ptr(i), max_timesteps, _fp(obs[i]), _fp(rew[i:i + 1]), _u8p(done[i:i + 1]))
        return obs, rew, done

    def get_obs(self):
        obs = np.zeros((self.n, OBS), np.float32)
        for i in range(self.n):
            lib().wo_env_get_obs(self.ptr(i), _fp(obs[i]))
        return obs

    def step(self, actions: np.ndarray, dt: float = DT_FRAME, iterations: int = 50, max_timesteps: int = 1000,
             auto_reset: bool = True, nthreads: int = 0):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, ACT)
        obs = np.zeros((self.n, OBS), np.float32)
        rew = np.zeros(self.n, np.float32)
        done = np.zeros(self.n, np.uint8)
        lib().wo_batch_step(self.base, self.n, _fp(a), dt, iterations, max_timesteps, int(auto_reset), _fp(obs), _fp(rew),
                            _u8p(done), nthreads)
        return obs, rew, done

    def get_state(self):
        """-> (f[n,92], i[n,2]) in the canonical record layout (walker_oracle.h)."""
        f = np.zeros((self.n, STATE_FLOATS), np.float32)
        iv = np.zeros((self.n, STATE_INTS), np.int32)
        for i in range(self.n):
            lib().wo_env_get_state(self.ptr(i), _fp(f[i]), _ip(iv[i]))
        return f, iv

    def set_state(self, f: np.ndarray, iv: np.ndarray):
        f = np.ascontiguousarray(f, np.float32).reshape(self.n, STATE_FLOATS)
        iv = np.ascontiguousarray(iv, np.int32).reshape(self.n, STATE_INTS)
        for i in range(self.n):
            lib().wo_env_set_state(self.ptr(i), _fp(f[i]), _ip(iv[i]))


def _is_triple(x) -> bool:
    return len(x) == 3 and all(isinstance(v, (float, int, np.floating)) for v in x)


def sat(a, b, ca=None, cb=None):
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 2)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 2)
    ca = np.ascontiguousarray(a.mean(0) if ca is None else ca, np.float32)
    cb = np.ascontiguousarray(b.mean(0) if cb is None else cb, np.float32)
    n = np.zeros(2, np.float32)
    d = np.zeros(1, np.float32)
    ax = np.zeros(1, np.int32)
    r = lib().wo_sat(_fp(a), len(a), _fp(b), len(b), _fp(ca), _fp(cb), _fp(n), _fp(d), _ip(ax))
    return bool(r), n, float(d[0]), int(ax[0])


def contacts(a, b, normal):
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 2)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 2)
    nrm = np.ascontiguousarray(normal, np.float32)
    pts = np.zeros(4, np.float32)
    k = lib().wo_contacts(_fp(a), len(a), _fp(b), len(b), _fp(nrm), _fp(pts))
    return pts.reshape(2, 2)[:k].copy()


# ------------------------------------------------------------------ PPO ----
DENSE, RELU, LEAKYRELU, TANH = 0, 1, 2, 3
ACTOR_LAYERS = [(DENSE, 64), (LEAKYRELU, 0), (DENSE, 64), (LEAKYRELU, 0), (DENSE, 4), (TANH, 0)]  # Hyperparameters.cs:92
CRITIC_LAYERS = [(DENSE, 64), (LEAKYRELU, 0), (DENSE, 1)]  # Hyperparameters.cs:91


def hyper_defaults() -> Hyper:
    hp = Hyper()
    lib().wo_hyper_defaults(C.byref(hp))
    return hp


class Net:
    def __init__(self, input_size: int, layers):
        kinds = np.array([k for k, _ in layers], np.int32)
        sizes = np.array([s for _, s in layers], np.int32)
        self.h = lib().wo_net_create(input_size, _ip(kinds), _ip(sizes), len(layers))
        self.input_size = input_size
        self.n_dense = int((kinds == DENSE).sum())
        self.num_params = lib().wo_net_num_params(self.h)
        self.out = lib().wo_net_output_size(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().wo_net_destroy(self.h)
            self.h = None

    def set_params(self, flat):
        flat = np.ascontiguousarray(flat, np.float32)
        assert flat.size == self.num_params
        lib().wo_net_set_params(self.h, _fp(flat))

    def get_params(self):
        out = np.zeros(self.num_params, np.float32)
        lib().wo_net_get_params(self.h, _fp(out))
        return out

    def get_grads(self):
        out = np.zeros(self.num_params, np.float32)
        lib().wo_net_get_grads(self.h, _fp(out))
        return out

    def get_adam(self):
        m = np.zeros(self.num_params, np.float32)
        v = np.zeros(self.num_params, np.float32)
        it = np.zeros(self.n_dense, np.int32)
        lib().wo_net_get_adam(self.h, _fp(m), _fp(v), _ip(it))
        return m, v, it

    def set_adam(self, m, v, it):
        lib().wo_net_set_adam(self.h, _fp(np.ascontiguousarray(m, np.float32)), _fp(np.ascontiguousarray(v, np.float32)),
                              _ip(np.ascontiguousarray(it, np.int32)))

    def forward(self, x, cache=False):
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim == 1:
            y = np.zeros(self.out, np.float32)
            lib().wo_net_forward(self.h, _fp(x), _fp(y), int(cache))
            return y
        y = np.zeros((x.shape[0], self.out), np.float32)
        for i in range(x.shape[0]):
            lib().wo_net_forward(self.h, _fp(x[i]), _fp(y[i]), int(cache))
        return y

    def optimise(self, hp: Hyper):
        lib().wo_net_optimise(self.h, C.byref(hp))


def ppo_train_batch(actor: Net, critic: Net, hp: Hyper, states, actions, old_logp, adv, ret, optimise=True):
    states = np.ascontiguousarray(states, np.float32)
    actions = np.ascontiguousarray(actions, np.float32)
    old_logp = np.ascontiguousarray(old_logp, np.float32)
    adv = np.ascontiguousarray(adv, np.float32)
    ret = np.ascontiguousarray(ret, np.float32)
    cl = np.zeros(1, np.float32)
    al = np.zeros(1, np.float32)
    skipped = lib().wo_ppo_train_batch(actor.h, critic.h, C.byref(hp), states.shape[0], _fp(states), _fp(actions),
                                       _fp(old_logp), _fp(adv), _fp(ret), int(optimise), _fp(cl), _fp(al))
    return skipped, float(cl[0]), float(al[0])


def sample_actions(actor: Net, hp: Hyper, state, u):
    state = np.ascontiguousarray(state, np.float32)
    u = np.ascontiguousarray(u, np.float32)
    a = np.zeros(actor.out, np.float32)
    lp = np.zeros(actor.out, np.float32)
    mu = np.zeros(actor.out, np.float32)
    lib().wo_sample_actions(actor.h, C.byref(hp), _fp(state), _fp(u), _fp(a), _fp(lp), _fp(mu))
    return a, lp, mu


def mc_returns(rewards, values, gamma):
    r = np.ascontiguousarray(rewards, np.float32)
    v = np.ascontiguousarray(values, np.float32)
    G = np.zeros_like(r)
    A = np.zeros_like(r)
    lib().wo_mc_returns(_fp(r), _fp(v), r.size, gamma, _fp(G), _fp(A))
    return G, A


def gae(rewards, values, gamma, lam):
    r = np.ascontiguousarray(rewards, np.float32)
    v = np.ascontiguousarray(values, np.float32)
    G = np.zeros_like(r)
    A = np.zeros_like(r)
    lib().wo_gae(_fp(r), _fp(v), r.size, gamma, lam, _fp(G), _fp(A))
    return G, A


def normalize(x, epsilon):
    x = np.array(x, np.float32, copy=True)
    lib().wo_normalize(_fp(x), x.size, epsilon)
    return x
