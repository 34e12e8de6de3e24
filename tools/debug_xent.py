"""Debug aid: the fused postprocess2 + cross-entropy kernel against an fp64 reference over a few shapes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
from wavenet._lib import load
lib = load()
p = lambda t: t.data_ptr()
Q = 256
for (B, T, K, wb) in [(1, 1000, 512, True), (1, 1000, 512, False), (1, 1000, 128, True), (1, 1000, 256, True), (2, 500, 512, True),
                      (1, 128, 512, True), (1, 1000, 192, True)]:
    rng = np.random.default_rng(K + T)
    M = B * T
    a16 = torch.tensor(np.maximum(rng.standard_normal((M, K)), 0).astype(np.float32)).half()
    w16 = torch.tensor((rng.standard_normal((Q, K)) * (2.0 / np.sqrt(K))).astype(np.float32)).half()
    bias = torch.tensor((0.5 * rng.standard_normal(Q)).astype(np.float32))
    ids = torch.tensor(rng.integers(0, Q, (B, T)).astype(np.int32))
    logits = a16.double() @ w16.double().T + (bias.double() if wb else 0.0)
    lse = torch.logsumexp(logits, 1)
    tgt = torch.cat([ids[:, 1:], torch.zeros(B, 1, dtype=torch.int32)], 1).reshape(M).long()
    valid = torch.ones(B, T, dtype=torch.bool); valid[:, -1] = False; valid = valid.reshape(M)
    row = (lse - logits[torch.arange(M), tgt]) * valid
    loss_ref = float(row.sum() / M)
    sm = torch.softmax(logits, 1)
    sm[torch.arange(M)[valid], tgt[valid]] -= 1.0
    da, dw, db, di = a16.cuda(), w16.cuda(), bias.cuda(), ids.cuda()
    partials = torch.zeros(4096, device='cuda'); out = torch.zeros((), device='cuda')
    g16 = torch.zeros(M, Q, dtype=torch.float16, device='cuda'); bg = torch.zeros(Q, device='cuda')
    rc = lib.wn_post2_xent(p(da), K, p(dw), K, p(db) if wb else None, p(di), B, T, K, Q, p(partials), p(out), p(g16),
                           1.5, p(bg), 1 / 1.5, None)
    torch.cuda.synchronize()
    g = g16.double().cpu() / 1.5
    err_rows = ((g - sm).norm(dim=1) / sm.norm(dim=1))
    print((B, T, K, wb), 'rc', rc, 'loss', float(out), 'ref', loss_ref, 'grad rel', float((g - sm).norm() / sm.norm()),
          'worst rows', err_rows.topk(3).indices.tolist(), err_rows.topk(3).values.tolist(),
          'bias rel', float((bg.double().cpu() - sm.sum(0)).norm() / sm.sum(0).norm()))
