"""Data-parallel diagnostic: do the ranks hold bit-identical gradients / parameters after every TrainStep?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_diag.py

For each configuration (eager / graph-captured step, gradient sink on / off) every rank runs the same steps on its
slice of a global batch; after each step the flat gradient buffer (post all-reduce) and the flat parameter buffer are
all-gathered and rank 0 prints which parameters differ between ranks and by how much."""
import os
import sys
import threading

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalaggressionrecognition_b200 import models as M, ops, training, workloads as W  # noqa: E402

KW = dict(t_audio=24, t_video=8)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    steps = int(os.environ.get("DP_DIAG_STEPS", "7"))
    precision = os.environ.get("DP_DIAG_PRECISION", "fp32")
    configs = ((False, "w,b"), (False, ""), (True, "w,b"), (True, ""))
    trace = os.environ.get("DP_DIAG_TRACE") == "1"
    events = []
    if trace:       # event log of one eager step with the gradient sink on: n<i> = notify, f<i> = hook fired, L<b> = bucket launched
        configs = ((False, "w,b"),)
        steps = 2
        orig_make, orig_notify, orig_launch = training.GradSync._make_hook, training.GradSync.notify, training.GradSync._launch

        def make_hook(self, i):
            inner = orig_make(self, i)

            def hook(p):
                events.append(f"f{i}")
                return inner(p)
            return hook

        def notify(self, param):
            events.append(f"n{self.index_of.get(id(param))}")
            return orig_notify(self, param)

        def launch(self, b):
            events.append(f"L{b}[{self.pending}]")
            return orig_launch(self, b)
        training.GradSync._make_hook, training.GradSync.notify, training.GradSync._launch = make_hook, notify, launch
    for graph, sink in configs:
        ops._SINK_KINDS = set(sink.split(",")) if sink else set()
        torch.manual_seed(0)
        model = W.perturb_norms(W.disable_dropout(W.build_c3(M, **KW))).to(dev).train()
        names = [n for n, p in model.named_parameters() if p.requires_grad]
        crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
        step = training.TrainStep(model, crit, lr=1e-3, graph=graph, precision=precision)
        if rank == 0:
            print(f"== graph={graph} sink='{sink}' buckets={[(lo, hi) for lo, hi, _, _ in step.sync.buckets]}", flush=True)
        for s in range(steps):
            data, labels = W.batch_c3(B=8 * world, seed=500 + s, **KW)
            sl = slice(rank * 8, (rank + 1) * 8)
            d = [[n[sl], t[sl]] for n, t in data]
            l = [[n[sl], y[sl]] for n, y in labels]
            events.clear()
            step(W.to_device(d, dev), W.to_device(l, dev))
            torch.cuda.synchronize()
            if trace:
                ev = [None] * world
                dist.all_gather_object(ev, list(events))
                if rank == 0:
                    for r, e in enumerate(ev):
                        print(f"  step {s} rank {r} events: " + " ".join(e), flush=True)
            for what, buf in (("grad", step.flat.grad), ("param", step.flat.flat)):
                mine = buf.detach().clone()
                allb = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(allb, mine)
                if rank == 0:
                    diff = (allb[0] - allb[1]).abs()
                    bad = []
                    for i, (p, o) in enumerate(zip(step.flat.params, step.flat.offsets)):
                        dmax = float(diff[o:o + p.numel()].max())
                        if dmax > 0:
                            ref = float(allb[0][o:o + p.numel()].abs().max())
                            bad.append(f"{i}:{names[i].replace('transformer_squence_processing', 'tsp').replace('modality_', 'm_')}:{dmax:.1e}/{ref:.1e}")
                    print(f"  step {s} {what}: {len(bad)} of {len(names)} tensors differ " + " ".join(bad[:12]), flush=True)
        step.release_graphs()
        del step, model
    sys.stdout.flush()
    t = threading.Timer(30.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
