import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from gaml_b200 import api, synth
wl = synth.paired_workload(46, 10000, 200_000, n_evals=6, seed=11)
whole = api.ProbCalculator.from_workload(wl)
ranks = [api.ProbCalculator.from_workload(wl, shard_of=(rk, 2)) for rk in range(2)]
ptrs = [pc.peer_exchange_create(rk, 2)[1] for rk, pc in enumerate(ranks)]
for pc in ranks:
    pc.peer_exchange_open(local_ptrs=ptrs)
for e, walks in enumerate(wl.evals):
    ref = whole.calc_prob(walks)
    for rk, pc in enumerate(ranks):
        t0 = time.time(); pc.prepare(walks); t1 = time.time(); pc.launch(); t2 = time.time()
        print(e, 'rank', rk, 'prepare %.3f ms launch %.3f ms' % ((t1-t0)*1e3, (t2-t1)*1e3), flush=True)
    for rk, pc in enumerate(ranks):
        t0 = time.time()
        try:
            g, tl = pc.finish_gathered()
            print(e, 'rank', rk, 'finish %.3f ms' % ((time.time()-t0)*1e3), pc.combine(g, 2, tl) == ref, flush=True)
        except Exception as ex:
            print(e, 'rank', rk, 'finish FAILED after %.3f ms' % ((time.time()-t0)*1e3), ex, flush=True)
