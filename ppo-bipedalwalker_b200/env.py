"""Host-side mirror of the reference's environment surface over libwalker_b200 (batched: N walkers in lockstep).

Reference classes mirrored (same method names and argument meaning; arrays replace the reference's Matrix[n,1]):
  IMaterial + Ice..SuperRubber   Materials/*.cs
  Environment                    Environment.cs:64 Update, :126 StepObjects, :176 InitialState
  Walker                         Walker/Walker.cs:66 TakeActions, :132 GetState, :108 GetPosition, :212 Reset
EnvBatch is the thin handle wrapper the two are built on.  Errors: the reference logs and continues
(RigidBody.cs:91-94); here a non-zero status raises WalkerB200Error carrying wb_last_error().
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import ACT, OBS, PAIR_SLOTS, STATE_FLOATS, STATE_INTS, Hyperparams, check, lib, ptr

DT_FRAME = float(np.float32(0.0166667))  # MonoGame fixed step: (float)TimeSpan.FromTicks(166667).TotalSeconds (Game1.cs:60,73)

PAIR_TRACE_DTYPE = np.dtype(
    [("other", "<i4"), ("aabb", "<i4"), ("sat", "<i4"), ("axis", "<i4"), ("nx", "<f4"), ("ny", "<f4"), ("depth", "<f4"),
     ("ncontacts", "<i4"), ("c0x", "<f4"), ("c0y", "<f4"), ("c1x", "<f4"), ("c1y", "<f4")]
)
JOINT_TRACE_DTYPE = np.dtype([("active", "<i4"), ("depth", "<f4")])


@dataclass(frozen=True)
class IMaterial:
    """Materials/IMaterial.cs:6-11 (Color is UI-only and omitted)."""
    InverseMass: float
    Restitution: float
    Friction: float
    id: int = -1

    def register(self) -> "IMaterial":
        """Make a user-defined material known to the library (the IMaterial plugin point)."""
        if self.id >= 0:
            return self
        out = C.c_int32(-1)
        check(lib().wb_material_register(self.InverseMass, self.Restitution, self.Friction, C.byref(out)))
        return IMaterial(self.InverseMass, self.Restitution, self.Friction, out.value)


Ice = IMaterial(11, 0.3, 0.0, 0)
Wood = IMaterial(20, 0.3, 0.01, 1)
Paper = IMaterial(1, 0.3, 0.1, 2)
Titanium = IMaterial(0.01, 0.1, 0.2, 3)
Carpet = IMaterial(5, 0.3, 0.8, 4)
Rubber = IMaterial(11, 0.7, 0.5, 5)
Metal = IMaterial(15, 0.3, 1.0, 6)
SuperRubber = IMaterial(11, 1.0, 1.0, 7)
MATERIALS = {"Ice": Ice, "Wood": Wood, "Paper": Paper, "Titanium": Titanium, "Carpet": Carpet, "Rubber": Rubber,
             "Metal": Metal, "SuperRubber": SuperRubber}


def default_hyperparams() -> Hyperparams:
    hp = Hyperparams()
    check(lib().wb_hyperparams_default(C.byref(hp)))
    return hp


def init(device: int = 0) -> None:
    check(lib().wb_init(device))


def _material_ids(spec, n: int):
    if spec is None:
        return None
    if isinstance(spec, IMaterial):
        spec = spec.register().id
    if isinstance(spec, str):
        spec = MATERIALS[spec].id
    if isinstance(spec, (int, np.integer)):
        return np.full(n, int(spec), np.uint8)
    arr = [(_material_ids(s, 1)[0] if not isinstance(s, (int, np.integer)) else int(s)) for s in spec]
    assert len(arr) == n
    return np.asarray(arr, np.uint8)


def pin_host(array: np.ndarray) -> np.ndarray:
    """Page-lock a C-contiguous numpy array in place (wb_host_pin) so that EnvBatch.step takes the zero-copy path for it.
    Call unpin_host before the array is freed."""
    assert array.flags["C_CONTIGUOUS"]
    check(lib().wb_host_pin(ptr(array), array.nbytes))
    return array


def unpin_host(array: np.ndarray) -> None:
    check(lib().wb_host_unpin(ptr(array)))


class EnvBatch:
    """N independent environments resident on the GPU (wb_env_batch handle)."""

    def __init__(self, n: int, floor_materials=None, walker_materials=None, hp: Hyperparams | None = None, stream: int | None = None):
        self.n = int(n)
        self.hp = hp if hp is not None else default_hyperparams()
        fm = _material_ids(floor_materials, self.n)
        wm = _material_ids(walker_materials, self.n)
        h = C.c_void_p()
        check(lib().wb_env_create(self.n, ptr(fm), ptr(wm), C.byref(self.hp), C.byref(h)))
        self._h = h
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_h", None) and lib is not None:  # `lib` is None during interpreter shutdown
            lib().wb_env_destroy(self._h)
            self._h = None

    __del__ = close

    def set_stream(self, cuda_stream: int):
        check(lib().wb_env_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_variant(self, lanes_per_env: int):
        check(lib().wb_env_set_variant(self._h, lanes_per_env))

    def get_variant(self) -> int:
        """Lanes per environment of the physics kernel in use (chosen from the batch size unless set_variant was called)."""
        v = C.c_int32(0)
        check(lib().wb_env_get_variant(self._h, C.byref(v)))
        return int(v.value)

    def sync(self):
        check(lib().wb_env_sync(self._h))

    def launch_count(self) -> int:
        out = C.c_int64(0)
        check(lib().wb_env_launch_count(self._h, C.byref(out)))
        return out.value

    # -- state blobs ("identical start state" contract); record-major [n,92] / [n,2] on the Python side
    def get_state(self):
        f = np.empty((STATE_FLOATS, self.n), np.float32)
        iv = np.empty((STATE_INTS, self.n), np.int32)
        check(lib().wb_env_get_state(self._h, ptr(f), ptr(iv)))
        return np.ascontiguousarray(f.T), np.ascontiguousarray(iv.T)

    def set_state(self, f: np.ndarray, iv: np.ndarray):
        f = np.ascontiguousarray(np.asarray(f, np.float32).reshape(self.n, STATE_FLOATS).T)
        iv = np.ascontiguousarray(np.asarray(iv, np.int32).reshape(self.n, STATE_INTS).T)
        check(lib().wb_env_set_state(self._h, ptr(f), ptr(iv)))

    def reset(self, mask=None, first_episode: bool = False):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        check(lib().wb_env_reset(self._h, ptr(m), int(first_episode)))

    # -- granular reference calls
    def take_actions(self, actions):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n, ACT)
        check(lib().wb_env_take_actions(self._h, ptr(a)))

    def step_objects(self, delta_time: float = DT_FRAME):
        check(lib().wb_env_step_objects(self._h, delta_time))

    def debug_contacts(self, delta_time: float = DT_FRAME):
        it = self.hp.iterations
        pt = np.zeros((self.n, it, PAIR_SLOTS), PAIR_TRACE_DTYPE)
        jt = np.zeros((self.n, it, 4), JOINT_TRACE_DTYPE)
        check(lib().wb_env_debug_contacts(self._h, delta_time, ptr(pt), ptr(jt)))
        return pt, jt

    def observe(self):
        obs = np.empty((self.n, OBS), np.float32)
        rew = np.empty(self.n, np.float32)
        done = np.empty(self.n, np.uint8)
        check(lib().wb_env_observe(self._h, ptr(obs), ptr(rew), ptr(done)))
        return obs, rew, done

    def get_obs(self):
        obs = np.empty((self.n, OBS), np.float32)
        check(lib().wb_env_get_obs(self._h, ptr(obs)))
        return obs

    # -- fused env-step (one kernel launch)
    def step(self, actions, delta_time: float = DT_FRAME, auto_reset: bool = True, out=None):
        """actions: numpy [n,4] or a pinned torch tensor.  Returns (obs, reward, done) host arrays."""
        if hasattr(actions, "data_ptr"):
            a = actions
        else:
            a = np.ascontiguousarray(actions, np.float32).reshape(self.n, ACT)
        if out is None:
            out = (np.empty((self.n, OBS), np.float32), np.empty(self.n, np.float32), np.empty(self.n, np.uint8))
        obs, rew, done = out
        check(lib().wb_env_step(self._h, ptr(a), delta_time, int(auto_reset), ptr(obs), ptr(rew), ptr(done)))
        return obs, rew, done

    def step_dev(self, actions_dev, obs_dev, reward_dev, done_dev, delta_time: float = DT_FRAME, auto_reset: bool = True):
        """Device pointers (ints or torch CUDA tensors); enqueues on the handle's stream, no sync."""
        check(lib().wb_env_step_dev(self._h, ptr(actions_dev), delta_time, int(auto_reset), ptr(obs_dev), ptr(reward_dev),
                                    ptr(done_dev)))


class Walker:
    """Walker/Walker.cs surface over a batch (the PPO brain lives in ppo.PPOAgent)."""

    def __init__(self, batch: EnvBatch):
        self._b = batch

    def TakeActions(self, actions):  # Walker.cs:66-75 (the environment clips first, Environment.cs:78)
        self._b.take_actions(actions)

    def GetState(self):  # Walker.cs:132-152 -> [n,12]
        return self._b.get_obs()

    def GetPosition(self):  # Walker.cs:108-111: the Body centroid at the last Update
        f, _ = self._b.get_state()
        return f[:, 62:64].copy()

    @property
    def Terminal(self):  # Walker.cs:23
        _, iv = self._b.get_state()
        return (iv[:, 0] & _lib.FLAG_TERMINAL) != 0

    def Reset(self, mask=None):  # Walker.cs:212-223
        self._b.reset(mask)


class Environment:
    """Environment.cs surface for N walkers.  Update() takes the actions instead of sampling them, so a
    policy (ppo.PPOAgent.SampleActions) or a test can inject them."""

    def __init__(self, n: int = 1, floor: IMaterial | str | None = None, walker: IMaterial | str | None = None,
                 hp: Hyperparams | None = None):
        self.batch = EnvBatch(n, floor_materials=floor, walker_materials=walker, hp=hp)
        self.walker = Walker(self.batch)
        self._state = self.batch.get_obs()  # InitialState, Environment.cs:176-180

    def InitialState(self):
        self._state = self.batch.get_obs()
        return self._state

    def StepObjects(self, deltaTime: float = DT_FRAME):  # Environment.cs:126-143
        self.batch.step_objects(deltaTime)

    def Update(self, deltaTime: float, actions, auto_reset: bool = True):
        """Environment.Update (Environment.cs:64-92) minus the policy and the trainer: returns (state, reward, terminal)."""
        self._state, reward, terminal = self.batch.step(actions, deltaTime, auto_reset)
        return self._state, reward, terminal

    def Reset(self, mask=None):  # Environment.cs:167-173
        self.batch.reset(mask)
        return self.InitialState()
