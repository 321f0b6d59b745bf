import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from tests.test_field_host import limbs, unlimbs
from tests.util import FIELDS, zkb
z=zkb()
for name in ("kat124","bn254","p256full"):
    p=FIELDS[name]; b=z.GpuBackend(0); b.set_field(p); n=b.stats()["nlimb"]; R=1<<(32*n); Rinv=pow(R,-1,p)
    rng=np.random.default_rng(6)
    xs=[0,1,p-1]+[int.from_bytes(rng.bytes(4*n),"little")%p for _ in range(1000)]
    ys=[p-1,1,p-1]+[int.from_bytes(rng.bytes(4*n),"little")%p for _ in range(1000)]
    A,B=limbs(xs,n),limbs(ys,n)
    def cmp(op,want):
        got=unlimbs(b.debug_field_ops(op,A,B))
        bad=[i for i,(g,w) in enumerate(zip(got,want)) if g!=w]
        print(name,"op",op,"mismatches",len(bad), [(hex(xs[i]),hex(ys[i]),hex(got[i]),hex(want[i])) for i in bad[:2]])
    cmp(7,[(x*y)%R for x,y in zip(xs,ys)])
    cmp(8,[(x*y)>>(32*n) for x,y in zip(xs,ys)])
    cmp(6,[x*Rinv%p for x in xs])
    cmp(5,[x*y*Rinv%p for x,y in zip(xs,ys)])
    cmp(4,[(x+y)%p for x,y in zip(xs,ys)])
    cmp(3,[(3*x*y*Rinv+y)%p for x,y in zip(xs,ys)])
