"""16 more seeded random quad-list sequences (tests/test_cache.py::random_lists) through the DEVICE cache bookkeeping
against the reference's own GetHeightMapForQuad -- the same comparison as the test suite, more seeds.
    python tools/cache_random_seeds.py      (needs a GPU and oracle/_ref)"""
import sys; import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import planet_b200 as pb
from oracle.bindings import PortOracle, RefOracle
from test_cache import run_random
pb.init(0)
port=PortOracle(); ref=RefOracle(); bad=0
p=pb.fbm_params(1,0.5,pb.FAST)
for seed in range(200,216):
    cache=pb.HeightMapCache(32,1024,1499,extra_slots=4000)
    def plan(q,b):
        d,n=cache.frame_device(pb.quads_to_device(q),18,p,b)
        return d.cpu().numpy().view(pb.TEXRECT_DTYPE).reshape(-1),n,cache.count
    try: run_random(ref,port,plan,seed)
    except AssertionError as e: bad+=1; print('seed',seed,'FAILED',str(e)[:120])
    cache.close()
print('device seeds done, failures:',bad)
