"""Logger — mirror of src/utils/logger.jl:7-29.

make_logger(run_name; to_terminal, to_tensorboard, to_json) builds a tee of sinks and installs
it as the global logger. Records keep the reference's names and keys verbatim:
  "Episode Statistics":  episode_return episode_length global_step steps_per_sec log_step_increment (ppo.jl:157)
  "Training Statistics": loss pg_loss v_loss entropy_loss log_step_increment                       (ppo.jl:247)
TensorBoard semantics follow TensorBoardLogger.jl: scalars are tagged "<message>/<key>" and the
step advances by `log_step_increment`.
"""
import json
import os
import sys
import time

_global_logger = None


def _num(v):
    """a float the way json.dumps writes it"""
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "Infinity" if v > 0 else "-Infinity"
    return repr(v)


class NullSink:
    def write(self, message, step, kwargs):
        pass

    def close(self):
        pass


class ConsoleSink(NullSink):
    def write(self, message, step, kwargs):
        sys.stderr.write("[ Info: %s  %s\n" % (message, "  ".join("%s=%s" % kv for kv in kwargs.items())))


class JSONSink(NullSink):
    """FormatLogger(JSON(), "logs/<run>.json"; append=true)"""

    def __init__(self, path):
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        self.f = open(path, "a")

    def write(self, message, step, kwargs):
        self.f.write(json.dumps({"level": "info", "msg": message, "kwargs": kwargs}) + "\n")

    def close(self):
        self.f.close()


class TBSink(NullSink):
    """TBLogger("logs/<run>"): TensorBoard event files written by the native writer of libcleanrl_cuda.so (crl_tb_*,
    csrc/tblog.cu): one Event per record holding its "<message>/<key>" scalars at the logger's step. Building a
    protobuf per scalar in Python cost ~100 us each and made the logger the bottleneck of small-batch runs (17 records
    per PPO update). Without the built library: a scalars.jsonl with the same (tag, step, value) triples."""

    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self.tb = None
        self.f = None
        self._cache = {}
        try:
            import ctypes as C
            from . import _lib
            lib = _lib.load()
            h = C.c_void_p()
            _lib.check(lib.crl_tb_open(logdir.encode(), C.byref(h)))
            self.lib, self.tb, self._C = lib, h, C
        except Exception:
            self.f = open(os.path.join(logdir, "scalars.jsonl"), "a", buffering=1 << 20)

    def write(self, message, step, kwargs):
        if self.tb is not None:
            keys = tuple(k for k in kwargs if k != "log_step_increment")
            ent = self._cache.get((message, keys))
            if ent is None:
                blob = b"".join(("%s/%s" % (message, k)).encode() + b"\0" for k in keys)
                ent = self._cache[(message, keys)] = (blob, (self._C.c_double * len(keys))(), len(keys))
            blob, arr, n = ent
            for i, k in enumerate(keys):
                arr[i] = kwargs[k]
            self.lib.crl_tb_scalars(self.tb, time.time(), step, n, blob, arr)
            return
        # one formatted write per record; same bytes as json.dumps({"tag":..,"step":..,"value":..}) per scalar
        self.f.write("".join(['{"tag": "%s/%s", "step": %d, "value": %s}\n' % (message, k, step, _num(v))
                              for k, v in kwargs.items() if k != "log_step_increment"]))

    def close(self):
        if self.tb is not None:
            self.lib.crl_tb_close(self.tb)
            self.tb = None
        elif self.f is not None:
            self.f.close()


class TeeLogger:
    def __init__(self, sinks):
        self.sinks = sinks
        self.step = 0
        self.records = 0

    def info(self, message, **kwargs):
        """@info message key=value ... ; `log_step_increment` advances the TensorBoard step. A record without the key
        advances it by 1, TensorBoardLogger.jl's default `step_increment` (the DQN and A2C records, dqn.jl:82,116 and
        a2c.jl:100,106, carry none); the PPO records pass it explicitly, 0 included (ppo.jl:156-157,246-247)."""
        self.step += int(kwargs.get("log_step_increment", 1))
        self.records += 1
        for s in self.sinks:
            s.write(message, self.step, kwargs)

    def close(self):
        for s in self.sinks:
            s.close()


def make_logger(run_name, to_terminal=True, to_tensorboard=True, to_json=False, log_dir="logs"):
    """Creates a logger and sets it as the global logger (logger.jl:7-29)."""
    global _global_logger
    sinks = []
    if to_terminal:
        sinks.append(ConsoleSink())
    if to_tensorboard:
        sinks.append(TBSink(os.path.join(log_dir, run_name)))
    if to_json:
        sinks.append(JSONSink(os.path.join(log_dir, "%s.json" % run_name)))
    if not sinks:
        sinks.append(NullSink())
    _global_logger = TeeLogger(sinks)
    return _global_logger


def global_logger():
    return _global_logger
