"""Developer tool: where does the device likelihood of the n = 1000 composite test problem deviate from an
extended-precision (80-bit numpy longdouble, computed here on the CPU in about a minute) evaluation?  Compares K,
diag(L), z and the two likelihood terms.  Result on B200 (round 2): K mean relative error 7e-15, diag(L) <= 1.6e-13,
likelihood 4.4e-13 - the device sides with the 80-bit value; the multi-threaded CPU oracle of some pool boxes was 1.6e-10 off."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessfundamentals_b200 import engine as eng  # noqa: E402
from tests.test_gpu_multi import _problem  # noqa: E402

n = 1000
tree, hp, x, y = _problem(n, 1, 7 + n)


def truth():
    LD = np.longdouble
    xs, ys = x.astype(LD).reshape(-1), y.astype(LD).reshape(-1)
    l1, l2, p, c = [LD(v) for v in hp]
    d = xs[:, None] - xs[None, :]
    pi = LD("3.14159265358979323846264338327950288")
    se = np.exp(LD(-0.5) * (d * d) / (l1 * l1))
    per = np.exp(LD(-2) * np.sin(pi * (np.abs(d) / p)) ** 2 / (l2 * l2))
    K = (se + per) * ((xs - c)[:, None] * (xs - c)[None, :]) + LD(1e-2) * np.eye(n, dtype=LD)
    L = np.zeros((n, n), dtype=LD)
    for j in range(n):
        v = K[j:, j] - L[j:, :j] @ L[j, :j]
        L[j, j] = np.sqrt(v[0]); L[j + 1:, j] = v[1:] / L[j, j]
    z = np.zeros(n, dtype=LD)
    for i in range(n):
        z[i] = (ys[i] - L[i, :i] @ z[:i]) / L[i, i]
    return dict(K=K.astype(np.float64), Ldiag=np.diag(L).astype(np.float64), z=z.astype(np.float64), quad=float(z @ z),
                logdet=float(np.sum(np.log(np.diag(L)))))


T = truth()
prog = eng.DeviceProgram.get(tree, 1, False, 1)
for env in ("default",):
    plan = eng.Plan([prog], [n], want_grad=False)
    plan.set_data(0, torch.tensor(x), torch.tensor(y)); plan.set_hp(0, hp, 1e-2)
    plan.eval(eng.STAGE_ASSEMBLE); torch.cuda.synchronize()
    K = plan.lower_matrix(0).cpu().numpy()
    tril = np.tril(np.ones((n, n), dtype=bool))
    rel = np.abs(K - T["K"])[tril] / np.abs(T["K"])[tril]
    print("K: max rel err %.3e  mean rel err %.3e  max abs err %.3e  signed mean rel %.3e" %
          (rel.max(), rel.mean(), np.abs(K - T["K"])[tril].max(), ((K - T["K"])[tril] / T["K"][tril]).mean()))
    dd = np.diag(K) - np.diag(T["K"])
    print("K diag: max abs err %.3e  mean signed %.3e" % (np.abs(dd).max(), dd.mean()))
    yrow = plan.buffer(0, eng.BUF_A)[:n, n].cpu().numpy()   # carried row n = y^T
    print("carried y row: max abs diff to y %.3e" % np.abs(yrow - y.reshape(-1)).max())
    plan.eval(eng.STAGE_POTRF | eng.STAGE_NLL); torch.cuda.synchronize()
    L = plan.lower_matrix(0).cpu().numpy()
    ld = np.diag(L)
    rd = (ld - T["Ldiag"]) / T["Ldiag"]
    print("diag L: max rel err %.3e  signed mean %.3e  sum(log) dev %.6e vs truth %.6e -> diff %.3e" %
          (np.abs(rd).max(), rd.mean(), np.sum(np.log(ld)), float(T["logdet"]), np.sum(np.log(ld)) - float(T["logdet"])))
    for b in range(8):
        sl = slice(128 * b, min(n, 128 * (b + 1)))
        print("   block %d: signed mean rel err of L_ii %.3e   max %.3e" % (b, rd[sl].mean(), np.abs(rd[sl]).max()))
    z = plan.buffer(0, eng.BUF_Z).cpu().numpy()[:n]
    print("z: max abs err %.3e   z^T z dev %.10f truth %.10f diff %.3e" % (np.abs(z - T["z"]).max(), z @ z, float(T["quad"]),
                                                                      z @ z - float(T["quad"])))
    terms = plan.last_terms()
    print("device terms (quad, logdet):", terms, " truth:", float(T["quad"]), float(T["logdet"]))
    nll = plan.results()[0][0]
    print("device nll %.13f   truth %.13f   rel %.3e" % (nll, 0.5 * float(T["quad"]) + float(T["logdet"]) + 0.5 * n * np.log(2 * np.pi),
                                                      (nll - (0.5 * float(T["quad"]) + float(T["logdet"]) + 0.5 * n * np.log(2 * np.pi))) / nll))
