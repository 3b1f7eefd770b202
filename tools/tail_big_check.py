"""Debug / regression check (GPU box): a thin-plate matrix that is indefinite as a whole (244 trailing points), a failed
update that must leave the model intact, then fused and batched predictions against a numpy solve."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g
rng=np.random.default_rng(2066)
def cloud(rng, n):
    d = rng.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.where(rng.random(n) < 0.75, 1.0, 2.0) * rng.uniform(0.97, 1.03, n)
    P = d * r[:, None] * rng.uniform(0.6, 1.0, 3)
    y = np.where(r > 1.5, 1.0, 0.0) + 0.01 * rng.standard_normal(n)
    return P, y
n = int(rng.choice([1, 2, 3, 17, 127, 128, 129, 255, 300, 511, 640, 1000, 1337, 1500]))
kind = str(rng.choice(["thin_plate", "gaussian", "laplace"]))
P, y = cloud(rng, n); p0 = 4.2 + rng.random(); s2 = np.full(n, 1e-3)
ctx=g.Context(); reg=g.GPRegressor("thin_plate", p0, ctx=ctx)
m=reg.create(P[:,0],P[:,1],P[:,2],y,s2)
print('n', n, 'tail', m.n_tail)
d = np.sqrt(((P[:,None,:]-P[None,:,:])**2).sum(-1)); K = 2*d**3 - 3*p0*d**2 + p0**3 + np.diag(s2)
a = np.linalg.solve(K,y)
def check(tag):
    for q in (8, 63, 130):
        Q = np.random.default_rng(q).uniform(-1.2,1.2,(q,3))
        dq = np.sqrt(((Q[:,None,:]-P[None,:,:])**2).sum(-1)); Kq = 2*dq**3 - 3*p0*dq**2 + p0**3
        f_t = Kq@a; v_t = p0**3 - np.einsum('ij,ij->i', Kq, np.linalg.solve(K,Kq.T).T)
        w = (-6*(p0-dq))*a[None,:]
        g_t = np.stack([(w*(Q[:,None,c]-P[None,:,c])).sum(1) for c in range(3)],1)
        f,v,gr,tx,ty = reg.evaluate(m,Q[:,0],Q[:,1],Q[:,2],var=True,tangent=True)
        f1 = reg.evaluate(m,Q[:,0],Q[:,1],Q[:,2])
        print(tag, 'q=%d'%q, 'alpha', np.abs(m.alpha-a).max()/np.abs(a).max(), 'f', np.abs(f-f_t).max(), 'f1', np.abs(f1-f_t).max(), 'v', np.abs(v-v_t).max(), 'grad', np.abs(gr-g_t).max()/np.abs(g_t).max(), flush=True)
check('fresh')
Pn, yn = cloud(np.random.default_rng(1), 69)
try:
    reg.update(m, Pn[:,0],Pn[:,1],Pn[:,2], yn, np.full(69, 0.1))
    print('update succeeded', m.n, m.n_tail)
except g.GPRegressionException as e:
    print('update failed as documented:', e.code, m.n, m.n_tail)
    check('after failed update')
