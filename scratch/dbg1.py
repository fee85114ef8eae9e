import sys, numpy as np
sys.path.insert(0, '.')
from gtsam_b200 import capi, datasets, problem as P
from oracle import oracle_py as O
ctx = capi.Context(0)
prob = datasets.make("bal_tiny", ncams=24, npoints=3000, visibility="scattered")
dev, orc = capi.DeviceProblem(ctx, prob), O.OracleProblem(prob)
dev.linearize(); orc.linearize()
lam = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
st, e0, e1, fv = dev.solve(lam); so, f0, f1, fo = orc.solve(lam)
print("status", st, so, "failvar", fv, fo)
fp, fvv, sp, sv, par = dev.cliques()
info = dev.symbolic_info()
dims = np.asarray(P.VAR_DIM)[prob.var_type]
print("ncliques", info.ncliques, "levels", info.nlevels)
bad = 0
for c in range(info.ncliques):
    a, b = dev.conditional(c), orc.conditional(c)
    err = np.abs(a-b).max() / max(1, np.abs(b).max())
    f = int(dims[fvv[fp[c]:fp[c+1]]].sum()); s = int(dims[sv[sp[c]:sp[c+1]]].sum())
    if c >= 3000 or err > 1e-7:
        print(c, "f", f, "s", s, "parent", par[c], "err", err, "nan" if np.isnan(a).any() else "")
        if err > 1e-7:
            bad += 1
            d = np.abs(a-b); i,j = np.unravel_index(np.nanargmax(d), d.shape); print("   worst at", i, j, a[i,j], b[i,j])
            if bad > 3: break
