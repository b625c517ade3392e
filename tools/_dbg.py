import os, sys, time
sys.path.insert(0, '.')
import numpy as np
import bench
from gaml_b200 import api
wl, shard = bench.make_workload(1, 0, 202, kind="c4shard")
pc, _, _ = bench.load_calculator(wl, shard, 0)
flat = [api.FlatWalks(w) for w in wl.evals]
pc.calc_prob_partial_flat(flat[0])
slow=[]
for e in range(1, 201):
    t0 = time.perf_counter()
    pc.calc_prob_partial_flat(flat[e])
    dt = (time.perf_counter() - t0) * 1e6
    st = pc.stats()
    if dt > 300 or e % 40 == 0: print(e, 'fast', st.fast_change_evals, 'e2e %.1f us prepare %.1f launch %.1f finish %.1f full %d touched %d appends %d rebuilds %d' % (dt, st.last_prepare_host_us, st.last_launch_host_us, st.last_finish_host_us, st.last_was_full, st.last_records_gathered, st.cache_appends, st.cache_rebuilds), flush=True)
