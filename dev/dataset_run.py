"""C5 on hardware (BASELINE configs[4]; FX/setup.cpp:5690-5753): the reference's own case driver (baseline/_ref/luw_reference_driver) on the staged dataset-generation
project, (a) the sequential loop of the reference on ONE GPU over the first `--seq-cases` angles, (b) latticeurbanwind_b200.dataset_replicas over all angles on
`--gpus` GPUs. Reports cases/hour of both and checks that the replicas write the same DG_<inflow>_<angle>_* files, byte for byte, as the sequential run.
usage: python dev/dataset_run.py --gpus 8 [--seq-cases 4] > gpurun_out/dataset.json"""
import argparse, hashlib, json, os, re, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from latticeurbanwind_b200 import dataset_replicas as R

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--seq-cases", type=int, default=4)
ap.add_argument("--steps", type=int, default=0, help="override the deck's run_nstep (0: keep)")
ap.add_argument("--stagger", type=float, default=1.1, help="seconds between replica starts (the driver's console log name has one-second resolution)")
a = ap.parse_args()
driver = os.path.join(ROOT, "baseline", "_ref", "luw_reference_driver")
src = os.path.join(ROOT, "baseline", "_ref", "case_dataset")
work = tempfile.mkdtemp(prefix="luw_dg_")
deck = open(os.path.join(src, "conf.luwdg")).read()
if a.steps > 0:
    deck = re.sub(r"(?m)^run_nstep\s*=.*$", f"run_nstep = {a.steps}", deck)
angles = R.parse_list(deck, "angle")


def outputs(project):
    found = {}
    for base, _, files in os.walk(project):
        for f in files:
            if f.startswith("DG_") or "/DG_" in os.path.join(base, f):
                p = os.path.join(base, f)
                found[os.path.relpath(p, project)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    return found


# (a) the reference's sequential loop, one process, one GPU
seq = os.path.join(work, "seq"); shutil.copytree(src, seq)
open(os.path.join(seq, "conf.luwdg"), "w").write(R._set_list(deck, "angle", angles[:a.seq_cases]))
t0 = time.time()
r = subprocess.run([driver, os.path.join(seq, "conf.luwdg")], cwd=seq, stdin=subprocess.DEVNULL, capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES="0"))
t_seq = time.time() - t0
open(os.path.join(ROOT, "gpurun_out", "dataset_sequential.log"), "w").write(r.stdout[-20000:] + "\n--- stderr\n" + r.stderr[-4000:])
assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
seq_out = outputs(seq)
# (b) replicas, one process per GPU
rep = os.path.join(work, "rep"); shutil.copytree(src, rep)
open(os.path.join(rep, "conf.luwdg"), "w").write(deck)
t0 = time.time()
res = R.launch(os.path.join(rep, "conf.luwdg"), driver, a.gpus, stagger_s=a.stagger)
t_rep = time.time() - t0
assert all(rc == 0 for _, _, rc, _ in res), [(d, rc, open(log).read()[-1500:]) for d, _, rc, log in res if rc != 0]
rep_out = outputs(rep)
tag = lambda name: re.search(r"DG_[^_]+_[^_]+_", name).group(0) if re.search(r"DG_[^_]+_[^_]+_", name) else name
common = sorted(set(seq_out) & set(rep_out))
missing = sorted(k for k in seq_out if k not in rep_out)
differing = [k for k in common if seq_out[k] != rep_out[k]]
line = {"cases": len(angles), "gpus": a.gpus, "sequential": {"cases": a.seq_cases, "seconds": t_seq, "cases_per_hour": a.seq_cases / t_seq * 3600.0},
        "replicas": {"cases": len(angles), "processes": len(res), "seconds": t_rep, "cases_per_hour": len(angles) / t_rep * 3600.0},
        "speedup_in_cases_per_hour": (len(angles) / t_rep) / (a.seq_cases / t_seq),
        "files": {"sequential": len(seq_out), "replicas": len(rep_out), "case_tags_replicas": len({tag(k) for k in rep_out}), "compared": len(common), "missing_in_replicas": missing[:8],
                  "byte_identical": len(common) - len(differing), "differing": differing[:8]},
        "stagger_s": a.stagger, "steps_per_case": a.steps if a.steps > 0 else 300,
        "deck": "baseline/_ref/case_dataset/conf.luwdg (example_DatasetGen: 16 inflow directions, 400 x 400 x 200 cells at 2.5 m, VTK output)"}
print(json.dumps(line))
shutil.rmtree(work, ignore_errors=True)
