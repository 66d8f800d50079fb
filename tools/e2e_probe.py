import sys, time, tempfile
sys.path.insert(0,'stress-particle-sph_b200')
import numpy as np, torch, spsph
from spsph import decks
sys.path.insert(0,'.')
from bench import pinned_like
d=tempfile.mkdtemp(); decks.write_deck(d, decks.refined_bui_spec(ncol=1632)); P=spsph.load(d,'bui')
eng=spsph.Engine(P)
pinned,keep=pinned_like(dict(P.arrays))
outp,keep2=pinned_like({k:P.arrays[k] for k in ("x","vel","stress","internal_vars","displ")})
for rep in range(3):
    t0=time.perf_counter(); eng.upload(pinned); eng.sync(); t1=time.perf_counter()
    eng.run(1,0.0,P.blocks[0]['dt'],3); t2=time.perf_counter()
    eng.download(outp); t3=time.perf_counter()
    print(f"upload {1e3*(t1-t0):.1f} ms  3 steps {1e3*(t2-t1):.1f} ms  download {1e3*(t3-t2):.1f} ms")
