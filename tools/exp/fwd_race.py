#!/usr/bin/env python
"""Run-to-run identity hunt: the module-path VideoNas forward (and backward) repeated on the same input; every output and
gradient of every repetition must be bit-identical to the first.  Prints which tensors / rows differ."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from computervision_codes_b200.tcn import VideoNas  # noqa: E402

DEV = "cuda"
C, D, T = 64, 2048, 1800
N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
CHURN = int(sys.argv[2]) if len(sys.argv) > 2 else 1
args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
torch.manual_seed(C)
m = VideoNas(args, 11, 10, 3, C, D, 100).to(DEV).eval()
x = torch.randn(1, T, D, device=DEV)
labels = [(torch.rand(T, k, device=DEV) < 0.05).float() for k in (6, 10, 15, 100)]
bce = torch.nn.BCEWithLogitsLoss()


def run():
    for p in m.parameters():
        p.grad = None
    outs = m(x, False)
    terms = [sum(bce(pd[0].transpose(0, 1), y) for pd in lst) for lst, y in zip((outs[1], outs[2], outs[3], outs[0]), labels)]
    loss = 0.1 * (terms[0] + terms[1] + terms[2]) + terms[3]
    loss.backward()
    flat = []
    for i, lst in enumerate(outs[:5]):
        for j, t in enumerate(lst):
            flat.append((f"out{i}.{j}", t.detach().clone()))
    for k, p in m.named_parameters():
        if p.grad is not None:
            flat.append(("grad:" + k, p.grad.detach().clone()))
    return flat


first = run()
nbad = 0
for it in range(N):
    if CHURN:  # disturb the caching allocator so recycled blocks carry other contents (NaN) into pad rows
        junk = [torch.full((int(torch.randint(1, 4000, (1,))), 64), float("nan"), device=DEV) for _ in range(20)]
        del junk
    cur = run()
    bad = [(n, float((a - b).abs().nan_to_num(1e9).max()), a, b) for (n, a), (_, b) in zip(cur, first) if not torch.equal(a, b)]
    if bad:
        nbad += 1
        print(f"iter {it}: {len(bad)} tensors differ; first 6:")
        for n, dmax, a, b in bad[:6]:
            dd = (a - b).abs().nan_to_num(1e9)
            if n.startswith("out"):
                rows = (dd.amax(dim=tuple(range(dd.dim() - 1))) > 0).nonzero().flatten()
                print(f"   {n} shape {tuple(a.shape)} max {dmax:.3e} frames differing {rows.numel()} "
                      f"[{int(rows.min())}..{int(rows.max())}]")
            else:
                print(f"   {n} max {dmax:.3e}")
print(f"{nbad} of {N} repetitions differ from the first")
