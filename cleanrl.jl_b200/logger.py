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

_global_logger = None


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
    """TBLogger("logs/<run>"): real TensorBoard event files when tensorboard is importable,
    else a scalars.jsonl with the same (tag, step, value) triples."""

    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self.writer = None
        try:
            from torch.utils.tensorboard import SummaryWriter  # needs the tensorboard package
            self.writer = SummaryWriter(logdir)
        except Exception:
            self.f = open(os.path.join(logdir, "scalars.jsonl"), "a")

    def write(self, message, step, kwargs):
        for k, v in kwargs.items():
            if k == "log_step_increment":
                continue
            tag = "%s/%s" % (message, k)
            if self.writer is not None:
                self.writer.add_scalar(tag, float(v), step)
            else:
                self.f.write(json.dumps({"tag": tag, "step": step, "value": float(v)}) + "\n")

    def close(self):
        if self.writer is not None:
            self.writer.close()
        else:
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
