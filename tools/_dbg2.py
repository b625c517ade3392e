import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GAML_B200_PATCH_DEBUG"] = "1"
from gaml_b200 import api, synth
wl = synth.paired_workload(10000, 10000, 200_000, n_evals=22, seed=44)
pc = api.ProbCalculator.from_workload(wl)
rev = list(reversed(wl.evals[0]))
script = [(0, True), (2, True), (0, True), (3, True), (4, False), (5, False), (6, True), (6, True), (1, True), (9, False),
          (0, True), (rev, True), (2, True), (12, True), (rev, True), (0, True), (21, True), (20, True)]
for step, (k, fresh) in enumerate(script):
    walks = wl.evals[k] if isinstance(k, int) else k
    if fresh:
        pc.reset_state()
    pc.calc_prob_partial(walks)
    st = pc.stats()
    print(step, k if isinstance(k, int) else "rev", fresh, len(walks), "patched", st.full_patch_evals, "reuse", st.full_reuse_evals, "h2d", st.last_h2d_bytes, flush=True)
