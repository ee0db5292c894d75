"""GPU diagnostic: per-parameter bf16 gradient error against the fp32 oracle for the golden cases."""
import sys, torch
sys.path.insert(0, ".")
import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import models as M, workloads as W
from oracle import oracle as O
from tests import helpers as H
O.DROPOUT_ENABLED = False
golden = torch.load("tests/golden/golden_v1.pt", weights_only=False)
dev = torch.device("cuda:0")
for name in ("c1_small", "c2_small", "c3_small"):
    spec = golden["cases"][name]["spec"]
    model, batch = H.build_case(spec, M, dev)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    data, labels = getattr(W, spec["batch"])(**spec["dkw"])
    po = H.oracle_forward(spec, sd, data, True, True)
    sum(H.oracle_losses(spec, po, labels).values()).backward()
    for eng in ("auto", "simt"):
        model.zero_grad()
        with mar.precision("bf16"), mar.engine(eng):
            model.train()
            _, losses = H.model_losses(spec, M, model, batch)
            losses.backward()
        got = {k: p.grad for k, p in model.named_parameters()}
        ref = {k: v.grad for k, v in sd.items()}
        tot = sum(float(r.double().pow(2).sum()) for r in ref.values() if r is not None)
        rows = []
        for k, r in ref.items():
            if r is None: continue
            d = float((got[k].double().cpu() - r.double()).pow(2).sum())
            rows.append((d / tot, H.rel_err(got[k], r), float(r.norm()), k))
        rows.sort(reverse=True)
        print(f"== {name} engine={eng}: whole-gradient rel err {H.global_rel_err(got, ref):.3e}")
        for share, rel, nrm, k in rows[:6]:
            print(f"   share of err^2 {share:.2e}  rel {rel:.3e}  |g| {nrm:.3e}  {k}")
