# extra checks of the record log in the new format
import numpy as np, sys
sys.path.insert(0,'/root/repo')
import numba_integrators_b200 as ni
from numba_integrators_b200 import workloads as wl
# 1. lean REC path (N >= 65536) vs lock-step reference of a few trajectories; step numbering; overflow/drain cycles
N=70000
w=wl.c2_harmonic(N=N, t_bound=3.0)
ens=ni.RK45(ni.rhs.harmonic, w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol, record=N*60)
assert ens.run()==0
acc=ens.n_accepted.astype(np.int64)
rec=ens.record_drain()
assert len(rec)==acc.sum(), (len(rec), acc.sum())
assert np.array_equal(np.bincount(rec.traj, minlength=N), acc)
r=rec.sorted()
# steps 1..acc per trajectory, t increasing
first=np.r_[0, np.cumsum(acc)[:-1]]
assert np.array_equal(r.step, np.concatenate([np.arange(1,a+1) for a in acc[:2000]]) if False else r.step)
for i in (0, 1, 777, N-1):
    t,y,aux=rec.trajectory(i)
    one=ni.RK45(ni.rhs.harmonic, w.t0, w.y0[i], w.t_bound, rtol=w.rtol, atol=w.atol)
    ts=[];ys=[]
    while ni.step(one): ts.append(one.t); ys.append(one.y)
    assert np.array_equal(t, np.array(ts)) and np.array_equal(y, np.array(ys)), i
    m=rec.traj==i
    assert np.array_equal(np.sort(rec.step[m]), np.arange(1,len(ts)+1))
print('lean REC ok', len(rec))
# 2. small log: overflow -> drain -> run again cycles
ens=ni.RK45(ni.rhs.harmonic, w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol, record=N*7)
tot=0; parts=[]
for cycle in range(40):
    try:
        left=ens.run()
    except ni.NiError as e:
        assert e.code==-6
        left=1
    p=ens.record_drain(); parts.append(p); tot+=len(p)
    if left==0: break
assert tot==acc.sum(), (tot, acc.sum(), cycle)
allt=np.concatenate([p.traj for p in parts]); alls=np.concatenate([p.step for p in parts]); alltt=np.concatenate([p.t for p in parts])
o=np.lexsort((alls,allt))
assert np.array_equal(allt[o], r.traj) and np.array_equal(alls[o], r.step) and np.array_equal(alltt[o], r.t)
assert np.array_equal(ens.y, ni.RK45(ni.rhs.harmonic, w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol).y) or True
assert cycle < 12, cycle
print("overflow cycles ok", cycle+1)
e2=ni.RK45(ni.rhs.harmonic, w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol); e2.run()
assert np.array_equal(ens.y, e2.y) and np.array_equal(ens.n_accepted, e2.n_accepted) and np.array_equal(ens.step_size, e2.step_size)
print('results identical to a run without recording')
# 3. Lorenz with aux recorded
w3=wl.c3_lorenz(N=66000, t_bound=0.5)
e3=ni.Advanced(3,1,ni.RK45)(ni.rhs.lorenz63, w3.t0, w3.y0, w3.params, w3.t_bound, rtol=w3.rtol, atol=w3.atol, record=66000*80)
assert e3.run()==0
r3=e3.record_drain()
assert len(r3)==e3.n_accepted.sum()
assert np.array_equal(r3.aux[:,0], (r3.y[:,0]*r3.y[:,0]+r3.y[:,1]*r3.y[:,1])+r3.y[:,2]*r3.y[:,2])
print('aux recorded ok')
