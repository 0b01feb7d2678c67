"""profiles/r02_fullsize.json: one summary row per named config at its stated size, from the bench lines kept in
profiles/ (bench.py attaches it as `extra.configs`).  usage: python scripts/r02/make_fullsize_summary.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = [("C2", "profiles/r02b_full_c2.json", "BASELINE.json configs[1]: 2-D flare current sheet, 4096^2, 1e8 particles"),
       ("C4", "profiles/r02c_full_c4.json", "configs[3]: PIC-derived diffusion + momentum diffusion, 4096^2, 1e7 particles"),
       ("C5", "profiles/r02i_full_c5.json", "configs[4]: 3-D flux rope, 512^3, 1.25e8 particles = the single-GPU share of 1e9 over 8 GPUs"),
       ("C5x8", "profiles/r02j_full_c5_n8.json", "configs[4] on 8 GPUs: 512^3 field per GPU, 1e9 particles in total")]
rows = []
for tag, path, what in SRC:
    p = os.path.join(ROOT, path)
    if not os.path.exists(p):
        continue
    d = json.loads(open(p).read().strip().splitlines()[-1])
    rows.append({
        "config": tag, "what": what, "source": path, "n_gpus": d["n_gpus"], "grid": d["config"]["grid"],
        "particles_per_gpu": d["config"]["particles_per_gpu"], "field_layout": d["config"]["field_layout"],
        "steps_per_s": d["value"], "e2e_steps_per_s": d["e2e"]["value"], "ms_per_step": d["ms_per_step"],
        "timed_steps": d["steps"], "warmup": d["warmup"],
        "algorithmic_gbs": d["roofline"]["achieved"], "frac_hbm": d["roofline"]["frac"], "clocks": d["clocks"],
    })
json.dump(rows, open(os.path.join(ROOT, "profiles", "r02_fullsize.json"), "w"), indent=1)
for r in rows:
    print(r["config"], "%.4g steps/s, e2e %.4g, %.0f ms/step, frac_hbm %.2f" % (r["steps_per_s"], r["e2e_steps_per_s"], r["ms_per_step"], r["frac_hbm"]))
