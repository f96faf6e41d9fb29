import sys, numpy as np
sys.path.insert(0,'/root/repo')
import newman_b200
from newman_b200 import pipeline, workloads
cfg = workloads.config("cfg3")
nr,nc,N = cfg["nr"],cfg["nc"],cfg["N"]
v = newman_b200.Mandelbrot(nr,nc,N=N,sz=cfg["sz"],center=cfg["center"],tol=cfg["tol"])
cand = [(r, c) for c in range(0,nc,2) for r in (nr//4, nr//2, 3*nr//4)] + [(r, nc//2) for r in range(0,nr,2)]
pix = np.array([r*nc+c for r,c in cand], dtype=np.int32)
uniq = np.unique(pix)
mk = lambda d: pipeline.TableSet(d, N, 1e-10, 1e-6, pipeline.floatexp_level(d))
dev = newman_b200.Device(0)
h0 = v.host_tables(nr//2, nc//2)
print("centre M", h0["M"])
# listed rounds
ts = mk(h0)
dev.frame_deep(ts.tables(), ts.arr["eps_re"], ts.arr["eps_im"], pix_list=uniq)
dev.launch()
gp, gi = dev.requeue()
print("glitched in round 0:", len(gp))
rnd=0
while len(gp):
    k = pipeline.pick_reference(gp, gi)
    ts = mk(v.host_tables(int(gp[k])//nc, int(gp[k])%nc))
    rnd+=1
    dev.frame_deep(ts.tables(), ts.arr["eps_re"], ts.arr["eps_im"], pix_list=np.ascontiguousarray(gp), mode=1 if rnd>=1 else 0)
    dev.launch()
    gp, gi = dev.requeue()
got = dev.read_pixels(pix)["iterations"]
order = np.argsort(-got, kind="stable")
print("top GPU counts:", [(cand[i], int(got[i])) for i in order[:10]])
iw = cand.index((6480,5760))
print("pinned winner GPU count", got[iw], "rank", int(np.where(order==iw)[0][0]))
for i in list(order[:6]) + [iw]:
    r,c = cand[i]
    print(cand[i], "gpu", int(got[i]), "exact", v.host_tables(r,c)["M"])
