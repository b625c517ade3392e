import os, sys, subprocess, tempfile
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from gaml_b200 import api, synth, workload
wl = synth.paired_workload(12, 2500, 4000, n_evals=30, seed=31)
with tempfile.TemporaryDirectory() as tmp:
    wp, rp = os.path.join(tmp, "a.wl"), os.path.join(tmp, "a.res")
    workload.write_workload(wp, wl)
    subprocess.run([os.path.abspath("oracle/gaml_oracle"), wp, rp, "1"], check=True, cwd=tmp, stderr=subprocess.DEVNULL)
    ref = workload.read_results(rp)
spec = wl.sets[0]
for mode in ("append", "noappend", "append_noperm"):
    os.environ.pop("GAML_B200_NO_APPEND", None); os.environ.pop("GAML_B200_NO_PERMUTE", None)
    if mode == "noappend": os.environ["GAML_B200_NO_APPEND"] = "1"
    if mode == "append_noperm": os.environ["GAML_B200_NO_PERMUTE"] = "1"
    pc = api.ProbCalculator(wl.node_len, wl.normalize_map)
    sid = pc.add_readset(spec)
    inserted = [set(), set()]
    for e, walks in enumerate(wl.evals):
        need = synth.short_keys_for_walks(walks, wl.node_len, with_single_node=True)
        nk = 0
        for m in range(2):
            for k in need:
                if k not in inserted[m] and k in spec.caches[m]:
                    pc.cache_insert(sid, m, k, spec.caches[m][k]); inserted[m].add(k); nk += 1
        prob, zeros, tl = pc.calc_prob(walks)
        v = pc.read_values(0)
        bad = np.nonzero(v != ref[e].per_read[0])[0]
        st = pc.stats()
        print(mode, e, 'newkeys', nk, 'full' if st.last_was_full else 'delta', 'bad', len(bad), bad[:5], 'zeros', zeros == ref[e].zeros, 'app', st.cache_appends, 'reb', st.cache_rebuilds)
        if len(bad):
            r = bad[0]; print('   read', r, v[r], ref[e].per_read[0][r])
            break
    pc.close()
