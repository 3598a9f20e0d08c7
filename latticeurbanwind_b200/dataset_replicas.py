"""Dataset generation as replicas: one case driver process per GPU.

The reference runs the `inflow` x `angle` cases of a `.luwdg` deck one after the other in one process (FX/setup.cpp:5690-5753); the cases are independent -- every
case builds its own LBM, voxelises, runs and writes `DG_<inflow>_<angle>_*` files (`:5740`) -- but the case driver keeps per-case state in process globals
(`units`, `coriolis_*_lbmu`, `buffer_*`, `sponge_*`: FX/setup.cpp:183-220, 5696-5716), so concurrency has to be process-level. This launcher splits the case list
into one Cartesian sub-deck per GPU, pins each process to its GPU (CUDA_VISIBLE_DEVICES, n_gpu = [1, 1, 1]) and runs them side by side. No collective, no
exchange: "replicas only".

Splitting rule. Before its case loop the driver derives a reference speed from the WHOLE list, si_ref_u = max(inflow) (FX/setup.cpp:3658); inside the loop every
case resets its units from its own inflow (`:5696-5703`). To stay on the safe side of that, the ANGLE list is split first and every sub-deck keeps the complete
inflow list, so that each process sees the same max(inflow) as the sequential run. Only when there are fewer angles than GPUs is the inflow list split too
(reported in the plan as `inflow_split`).

    python -m latticeurbanwind_b200.dataset_replicas <deck.luwdg> --driver baseline/_ref/luw_reference_driver --gpus 8
"""
import argparse
import os
import re
import subprocess
import sys

_LIST = r"(?m)^(\s*%s\s*=\s*)\[([^\]]*)\]\s*$"


def parse_list(deck, key):
    """The float list of `key = [a, b, ...]` as the deck spells it (strings are kept so that the sub-decks repeat the author's literals)."""
    m = re.search(_LIST % re.escape(key), deck)
    if not m:
        return []
    return [v.strip() for v in m.group(2).split(",") if v.strip()]


def _set_list(deck, key, values):
    text, n = re.subn(_LIST % re.escape(key), lambda m: m.group(1) + "[" + ", ".join(values) + "]", deck)
    if n != 1:
        raise ValueError(f"deck has no single `{key} = [...]` line")
    return text


def _chunks(items, n):
    """n contiguous chunks whose sizes differ by at most one (empty chunks dropped)."""
    n = max(1, min(n, len(items)))
    q, r = divmod(len(items), n)
    out, i = [], 0
    for k in range(n):
        size = q + (1 if k < r else 0)
        out.append(items[i:i + size])
        i += size
    return [c for c in out if c]


def plan(deck, gpus):
    """-> list of dicts {deck, inflow, angle, cases, inflow_split}: at most `gpus` Cartesian sub-decks that cover every (inflow, angle) pair exactly once."""
    inflow, angle = parse_list(deck, "inflow"), parse_list(deck, "angle")
    if not inflow or not angle:
        raise ValueError("a dataset-generation deck needs inflow = [...] and angle = [...] (FX/setup.cpp:3652, 5643)")
    gpus = max(1, int(gpus))
    angle_chunks = _chunks(angle, gpus)
    inflow_chunks = [inflow]
    if len(angle_chunks) < gpus and len(inflow) > 1:  # fewer angles than GPUs: split the inflows as well
        inflow_chunks = _chunks(inflow, gpus // len(angle_chunks))
    subs = []
    for ic in inflow_chunks:
        for ac in angle_chunks:
            text = _set_list(_set_list(deck, "inflow", ic), "angle", ac)
            text = re.sub(r"(?m)^(\s*n_gpu\s*=\s*).*$", r"\g<1>[1, 1, 1]", text)
            subs.append(dict(deck=text, inflow=ic, angle=ac, cases=[(i, a) for i in ic for a in ac], inflow_split=len(inflow_chunks) > 1))
    assert len(subs) <= gpus
    return subs


def launch(deck_path, driver, gpus, devices=None, dry_run=False, env=None):
    """Write one sub-deck per replica next to the original deck (the project's inputs are addressed relative to it), start one driver process per GPU with
    CUDA_VISIBLE_DEVICES set, wait for all. Returns [(device, sub-deck path, return code, log path)]. `devices`: CUDA ordinals to use (default 0..gpus-1)."""
    deck_path = os.path.abspath(deck_path)
    project = os.path.dirname(deck_path)
    stem, ext = os.path.splitext(os.path.basename(deck_path))
    subs = plan(open(deck_path).read(), gpus)
    devices = list(range(len(subs))) if devices is None else list(devices)
    if len(devices) < len(subs):
        raise ValueError("fewer devices than replicas")
    procs = []
    for k, sub in enumerate(subs):
        path = os.path.join(project, f"{stem}.replica{k}{ext}")
        with open(path, "w") as fh:
            fh.write(sub["deck"])
        log = os.path.join(project, f"{stem}.replica{k}.log")
        if dry_run:
            procs.append((devices[k], path, None, log, None))
            continue
        e = dict(os.environ if env is None else env, CUDA_VISIBLE_DEVICES=str(devices[k]))
        fh = open(log, "w")
        procs.append((devices[k], path, subprocess.Popen([driver, path], cwd=project, env=e, stdout=fh, stderr=subprocess.STDOUT), log, fh))
    out = []
    for dev, path, p, log, fh in procs:
        rc = None if p is None else p.wait()
        if fh is not None:
            fh.close()
        out.append((dev, path, rc, log))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("deck")
    ap.add_argument("--driver", required=True, help="the case driver binary (baseline/_ref/luw_reference_driver or a build of the reference tree against host/lbm.hpp)")
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--devices", default="", help="comma-separated CUDA ordinals (default 0..gpus-1)")
    ap.add_argument("--dry-run", action="store_true", help="write the sub-decks and print the plan, start nothing")
    a = ap.parse_args(argv)
    devices = [int(v) for v in a.devices.split(",")] if a.devices else None
    res = launch(a.deck, a.driver, a.gpus, devices, a.dry_run)
    for dev, path, rc, log in res:
        print(f"gpu {dev}: {os.path.basename(path)} -> {'planned' if rc is None else 'exit %d' % rc} ({log})")
    return 0 if all(rc in (None, 0) for _, _, rc, _ in res) else 1


if __name__ == "__main__":
    sys.exit(main())
