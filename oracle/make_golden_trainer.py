"""Golden vectors for the epoch bookkeeping of the reference's trainers.

    python oracle/make_golden_trainer.py     # writes tests/golden/golden_trainer_epoch.pt

Runs, in the build container, the LIVE reference `trainer.py` methods that turn per-step losses / logits / labels
into the per-epoch results dictionary — `nn_output_processing`, `compute_batch_loss`, `create_batch_results_dict`,
`compute_epoch_results` of `TorchSupervisedTrainer` (trainer.py:165-179, :235-286), `RNN_trainer` (:718-812) and
`MultimodalTrainer` (:888-1007) — on seeded synthetic step streams (tests/helpers.py::trainer_stream), and
records what they return.  tests/test_host_logic_cpu.py feeds the same streams to
`multimodalaggressionrecognition_b200.training.EpochAccumulator` (one device->host read per epoch instead of
2 x heads + 1 per step) and compares.  The reference's methods only read `self.metrics_dict`,
`self.train_samples_num` and `self.test_samples_num`, so they are called on a bare namespace object (the
constructors need datasets and a saving directory).  `matplotlib` is not installed and is only used by the
reference's plotting: a stub module stands in for it.
"""
from __future__ import annotations

import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
if "matplotlib" not in sys.modules:
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot

import trainer as ref_trainer  # noqa: E402  the reference

from tests import helpers as H  # noqa: E402


def run_reference(kind: str, stream, dataset_size: int):
    cls = {"single": ref_trainer.TorchSupervisedTrainer, "multi_head": ref_trainer.RNN_trainer,
           "multimodal": ref_trainer.MultimodalTrainer}[kind]
    me = types.SimpleNamespace(metrics_dict=H.trainer_metrics(), train_samples_num=dataset_size, test_samples_num=dataset_size)
    epoch = []
    for data, losses, pred, labels in stream:
        size = len(data[0]) if isinstance(data, list) else len(data)                 # trainer.py:155-158
        ret_loss = cls.compute_batch_loss(me, losses, size)
        pred_vals = cls.nn_output_processing(me, pred)
        epoch.append(cls.create_batch_results_dict(me, ret_loss, pred_vals, labels))
    return cls.compute_epoch_results(me, epoch, "train")


def main():
    import contextlib
    import io
    out = {"torch": torch.__version__, "cases": {}}
    for name, spec in H.TRAINER_STREAMS.items():
        with contextlib.redirect_stdout(io.StringIO()):      # MultimodalTrainer.compute_epoch_results prints its inputs (DEBUG)
            res = run_reference(spec["kind"], H.trainer_stream(**spec), spec["dataset_size"])
        out["cases"][name] = {"spec": spec, "results": res}
        print(name, res)
    path = os.path.join(ROOT, "tests", "golden", "golden_trainer_epoch.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
