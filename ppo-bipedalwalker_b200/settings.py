"""Hyper-parameter file interop: the JSON document the reference writes with
`Hyperparameters.SerializeJson` and reads with `DeserializeJson` (Walker/PPO/Hyperparameters.cs:124-187; property names of
`SerializableHyperparameters`, :11-49; range checks of `ValidateHyperparameterValues`, :189-217), so a settings file saved by
the original C# app configures this library and vice versa.  Host side only: nothing here touches the GPU.
"""
from __future__ import annotations

import json
import os
from dataclasses import asdict, dataclass, field

from ._lib import Hyperparams
from .ppo import DEFAULT_ACTOR, DEFAULT_CRITIC


@dataclass
class Settings:
    """`SerializableHyperparameters` with the defaults of Hyperparameters.cs:83-121 (same names, same order)."""
    GameSpeed: int = 1
    CollectData: bool = True
    SaveWeights: bool = True
    Iterations: int = 50
    MaxTimesteps: int = 1000
    RoughFloor: bool = False
    CriticNeuralNetwork: str = DEFAULT_CRITIC
    ActorNeuralNetwork: str = DEFAULT_ACTOR
    CriticWeightFileName: str = "critic"
    ActorWeightFileName: str = "actor"
    FilePath: str = field(default_factory=os.getcwd)
    Alpha: float = 0.001
    Beta1: float = 0.9
    Beta2: float = 0.999
    AdamEpsilon: float = 1e-8
    Epochs: int = 5
    BatchSize: int = 64
    UseGAE: bool = False
    NormalizeAdvantages: bool = False
    Gamma: float = 0.9
    Lambda: float = 0.95
    Epsilon: float = 0.3
    LogStandardDeviation: float = -1.0

    def validate(self) -> None:
        """ValidateHyperparameterValues (Hyperparameters.cs:189-217): same ranges, same (in)equalities."""
        def bad(cond, msg):
            if cond:
                raise ValueError(msg)
        bad(self.GameSpeed <= 0 or self.GameSpeed >= 10, f"Invalid game speed, should be in range 0<x<10 ({self.GameSpeed})")
        bad(self.Iterations <= 0 or self.Iterations >= 200, f"Invalid iterations count, should be in range 0<x<200 ({self.Iterations})")
        bad(self.MaxTimesteps <= 0, f"Invalid maximum time steps amount, should be in range x>0 ({self.MaxTimesteps})")
        bad(self.Alpha <= 0 or self.Alpha >= 10, f"Invalid alpha value, should be in range 0<x<10 ({self.Alpha})")
        bad(self.Beta1 <= 0 or self.Beta1 > 1, f"Invalid beta1 value, should be in range 0<x<1 ({self.Beta1})")
        bad(self.Beta2 <= 0 or self.Beta2 > 1, f"Invalid beta2 value, should be in range 0<x<1 ({self.Beta2})")
        bad(self.AdamEpsilon <= 0 or self.AdamEpsilon >= 1, f"Invalid Adam epsilon value, should be in range 0<x<1 ({self.AdamEpsilon})")
        bad(self.Epochs <= 0 or self.Epochs >= 50, f"Invalid epochs value, should be in range 0<x<50 ({self.Epochs})")
        bad(self.BatchSize <= 0 or self.BatchSize >= 1000, f"Invalid batch size value, should be in range 0<x<1000 ({self.BatchSize})")
        bad(self.Gamma <= 0 or self.Gamma > 1, f"Invalid gamma value, should be in range 0<x<1 ({self.Gamma})")
        bad(self.Lambda <= 0 or self.Lambda > 1, f"Invalid lambda value, should be in range 0<x<1 ({self.Lambda})")
        bad(self.Epsilon <= 0 or self.Epsilon > 1, f"Invalid epsilon value, should be in range 0<x<1 ({self.Epsilon})")
        bad(self.LogStandardDeviation <= -5 or self.LogStandardDeviation >= 5,
            f"Invalid log standard deviation value, should be in range -5<x<5 ({self.LogStandardDeviation})")

    def to_hyperparams(self) -> Hyperparams:
        """The plain struct the C ABI takes (wb_hyperparams): only the fields the hot path reads."""
        hp = Hyperparams()
        hp.iterations, hp.max_timesteps, hp.batch_size = self.Iterations, self.MaxTimesteps, self.BatchSize
        hp.use_gae, hp.normalize_advantages = int(self.UseGAE), int(self.NormalizeAdvantages)
        hp.alpha, hp.beta1, hp.beta2, hp.adam_epsilon = self.Alpha, self.Beta1, self.Beta2, self.AdamEpsilon
        hp.gamma, hp.lambda_, hp.epsilon, hp.log_std = self.Gamma, self.Lambda, self.Epsilon, self.LogStandardDeviation
        return hp


def SerializeJson(settings: Settings, fileLocation: str) -> None:
    """Hyperparameters.SerializeJson (:124-131): indented JSON, one property per SerializableHyperparameters member."""
    os.makedirs(os.path.dirname(os.path.abspath(fileLocation)), exist_ok=True)
    with open(fileLocation, "w") as fh:
        json.dump(asdict(settings), fh, indent=2)


def DeserializeJson(fileLocation: str, current: Settings | None = None, log=None) -> Settings:
    """Hyperparameters.DeserializeJson (:135-187).  Like the reference, a malformed document or an out-of-range value is
    LOGGED and the current settings are kept (the reference never surfaces these errors to its caller); a BatchSize <= 0 that
    slipped through becomes 64 (:168).  Missing properties keep System.Text.Json's defaults (0 / false / null) -- and therefore
    fail validation exactly as they do in the reference."""
    current = current if current is not None else Settings()
    log = log if log is not None else (lambda msg: None)
    try:
        with open(fileLocation) as fh:
            doc = json.load(fh)
        if not isinstance(doc, dict):
            raise ValueError("top-level JSON value is not an object")
    except Exception as exc:  # JsonSerializer.Deserialize threw
        log(f"JSON deserializer error. ({exc})")
        return current
    zero = {int: 0, bool: False, float: 0.0, str: None}
    fields = Settings.__dataclass_fields__
    try:
        values = {}
        for name, f in fields.items():
            kind = {"int": int, "bool": bool, "float": float, "str": str}[f.type if isinstance(f.type, str) else f.type.__name__]
            v = doc.get(name, zero[kind])
            if kind is float and isinstance(v, int) and not isinstance(v, bool):
                v = float(v)
            if v is not None and not isinstance(v, kind) or (kind is int and isinstance(v, bool)):
                raise ValueError(f"The JSON value could not be converted to {kind.__name__} ({name})")
            values[name] = v
        loaded = Settings(**values)
        loaded.validate()
        if loaded.BatchSize <= 0:
            loaded.BatchSize = 64
    except Exception as exc:
        log(f"Exception occurred while setting the values of the hyperparameters during deserialization: ({exc})")
        return current
    return loaded


__all__ = ["Settings", "SerializeJson", "DeserializeJson"]
