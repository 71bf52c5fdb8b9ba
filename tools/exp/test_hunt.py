"""The live-reference VideoNas test body (tests/test_full_size_gpu.py) with post-mortem localisation: when the first
forward of the drop-in differs from its second forward, walk both autograd graphs (no kernels added to the first run) and
name the first custom node whose saved input agrees while its saved h / masks / consumer input differ."""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
DEV = "cuda"


@pytest.fixture(autouse=True, scope="module")
def _flags():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    yield


def _nodes(outs):
    seen, order = set(), []

    def visit(fn):
        if fn is None or fn in seen:
            return
        seen.add(fn)
        for nxt, _ in fn.next_functions:
            visit(nxt)
        order.append(fn)  # post-order: producers first

    for lst in outs[:5]:
        for t in lst:
            visit(t.grad_fn)
    return [f for f in order if type(f).__name__ in ("DilatedResidualFnBackward", "TapGemmFnBackward")]


def test_hunt():
    from oracle import ref_import
    from computervision_codes_b200.tcn import VideoNas

    sys.setrecursionlimit(100000)
    C, D, T = 64, 2048, 1800
    net = ref_import.tenco_network()
    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(C)
    ref = net.VideoNas(args, 11, 10, 3, C, D, 100).to(DEV).eval()
    m = VideoNas(args, 11, 10, 3, C, D, 100).to(DEV).eval()
    m.load_state_dict(ref.state_dict())
    x = torch.randn(1, T, D, device=DEV)
    labels = [(torch.rand(T, k, device=DEV) < 0.05).float() for k in (6, 10, 15, 100)]
    bce = torch.nn.BCEWithLogitsLoss()

    def run(model, backward=True):
        outs = model(x, False)
        terms = [sum(bce(pd[0].transpose(0, 1), y) for pd in lst) for lst, y in zip((outs[1], outs[2], outs[3], outs[0]), labels)]
        loss = 0.1 * (terms[0] + terms[1] + terms[2]) + terms[3]
        if backward:
            loss.backward()
        return outs, loss

    if os.environ.get("HUNT_PRECREATE"):
        from computervision_codes_b200.layout import SeqLayout
        m._get_executor(SeqLayout.uniform(1, T, DEV))
        torch.cuda.synchronize()
    big = torch.randn(8192, 8192, device=DEV)
    import ctypes
    from computervision_codes_b200 import _lib
    from computervision_codes_b200.executor import _view

    def snapshot():
        ex = m._executor
        lib = _lib.load()
        snap = []
        for kind in (0, 1):
            for i in range(ex.cfg.layers_pg + 3 * ex.cfg.layers_r + (1 if kind == 0 else 0)):
                ptr = ctypes.c_void_p()
                assert lib.tcn_model_debug_ptr(ex.h, kind, i, ctypes.byref(ptr)) == 0
                snap.append((kind, i, _view(ptr.value, 1920, 64, DEV).clone()))
        return snap

    o_ref, _ = run(ref)
    if os.environ.get("HUNT_SYNC"):
        torch.cuda.synchronize()
    if os.environ.get("HUNT_BUSY"):   # a deterministic backlog: ~0.3 s of queued work ahead of the drop-in's first forward
        for _ in range(int(os.environ["HUNT_BUSY"])):
            big = (big @ big).clamp_(-1, 1)
    o1, _ = run(m, backward=False)
    s1 = snapshot()
    o2, _ = run(m, backward=False)
    s2 = snapshot()
    torch.cuda.synchronize()
    worst = max(float((a - b).abs().max()) for al, bl in zip(o1[:5], o2[:5]) for a, b in zip(al, bl))
    if worst <= 1e-6:
        return
    lines = [f"first forward differs from the second by {worst:.3e}"]
    shown = 0
    for (kind, i, a), (_, _, b) in zip(s1, s2):
        d = (a[:1800] - b[:1800]).abs().nan_to_num(1e9)
        if float(d.max()) > 0 and shown < 16:
            shown += 1
            rows = (d.amax(1) > 0).nonzero().flatten()
            cols = (d.amax(0) > 0).nonzero().flatten()
            blocks = sorted(set((rows // 128).tolist()))
            lines.append(f"  {'act' if kind == 0 else 'h'}[{i}]: max {float(d.max()):.3e} rows {rows.numel()} "
                         f"[{int(rows.min())}..{int(rows.max())}] blocks {blocks} cols {cols.numel()} "
                         f"[{int(cols.min())}..{int(cols.max())}]  first-run absmax {float(a[:1800].abs().max()):.3e} "
                         f"second-run absmax {float(b[:1800].abs().max()):.3e}")
    pytest.fail("\n".join(lines))
    n1, n2 = _nodes(o1), _nodes(o2)
    for i, (a, b) in enumerate(zip(n1, n2)):
        sa, sb = a.saved_tensors, b.saved_tensors
        dx = float((sa[0] - sb[0]).abs().nan_to_num(1e9).max())
        name = type(a).__name__
        if name.startswith("Dilated"):
            dh = float((sa[1] - sb[1]).abs().nan_to_num(1e9).max())
            dm = int((a.masks != b.masks).sum()) if a.masks is not None else -1
            if dh > 0 or dx > 0 or dm > 0:
                d = (sa[1] - sb[1]).abs().nan_to_num(1e9)
                rows = (d.amax(1) > 0).nonzero().flatten()
                cols = (d.amax(0) > 0).nonzero().flatten()
                rr = f"h rows {rows.numel()} [{int(rows.min())}..{int(rows.max())}] cols {cols.numel()} [{int(cols.min())}..{int(cols.max())}]" if rows.numel() else ""
                lines.append(f"  node {i} {name} shifts {a.shifts}: x diff {dx:.3e}  h diff {dh:.3e}  mask words differing {dm}  {rr}")
                shown += 1
        else:
            if dx > 0:
                d = (sa[0] - sb[0]).abs().nan_to_num(1e9)
                rows = (d.amax(1) > 0).nonzero().flatten()
                lines.append(f"  node {i} {name} w {tuple(sa[1].shape)}: x diff {dx:.3e} rows {rows.numel()} [{int(rows.min())}..{int(rows.max())}]")
                shown += 1
        if shown >= 14:
            break
    pytest.fail("\n".join(lines))
