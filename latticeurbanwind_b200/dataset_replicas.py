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

Grid. A sub-deck runs on ONE GPU (n_gpu = [1, 1, 1]). With mesh_control = "cell_size" that changes nothing. With mesh_control = "gpu_memory" the driver FITS the cell
size to a per-device memory budget (fit_cell_size_to_gpu_memory_request, FX/setup.cpp:371-405, through vram_required_mb_per_device(Nx, Ny, Nz, Dx, Dy, Dz)), so a
deck written for n_gpu = [2, 1, 1] would resolve to a coarser grid in every replica than in the sequential run -- a different dataset. plan() therefore refuses
such a deck unless it is told how to pin the grid: `cell_size` (the sequential run's cell size: the sub-decks get mesh_control = "cell_size"), or `regrid=True`
(gpu_memory is multiplied by Dx*Dy*Dz: close to, but not guaranteed to be, the sequential grid; reported as grid_pinned = False).

    python -m latticeurbanwind_b200.dataset_replicas <deck.luwdg> --driver baseline/_ref/luw_reference_driver --gpus 8 [--cell-size 4.0 | --regrid]
"""
import argparse
import os
import re
import subprocess
import sys
import time

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


def _scalar(deck, key):
    m = re.search(r"(?m)^\s*%s\s*=\s*(.*?)\s*$" % re.escape(key), deck)
    return m.group(1).strip().strip('"') if m else ""


def _set_scalar(deck, key, value):
    text, n = re.subn(r"(?m)^(\s*%s\s*=\s*).*$" % re.escape(key), lambda m: m.group(1) + value, deck)
    return text if n else text.rstrip("\n") + f"\n{key} = {value}\n"


def one_gpu_grid(deck, cell_size=None, regrid=False):
    """The deck rewritten for one GPU with the SAME grid as the sequential run -> (deck text, grid_pinned). See the module docstring (Grid)."""
    domains = 1
    for v in parse_list(deck, "n_gpu"):
        domains *= max(1, int(float(v)))
    text = re.sub(r"(?m)^(\s*n_gpu\s*=\s*).*$", r"\g<1>[1, 1, 1]", deck)
    if cell_size is not None:
        return _set_scalar(_set_scalar(text, "mesh_control", '"cell_size"'), "cell_size", repr(float(cell_size))), True
    if _scalar(deck, "mesh_control") != "gpu_memory" or domains == 1:
        return text, True
    if not regrid:
        raise ValueError(f'mesh_control = "gpu_memory" with n_gpu of {domains} devices: the driver fits the cell size to the PER-DEVICE budget (FX/setup.cpp:371-405), so one-GPU '
                         "replicas would compute a coarser grid than the sequential run. Pass cell_size=<the sequential run's cell size> to pin the grid, or regrid=True "
                         "to scale gpu_memory by the device count (approximate).")
    try:
        budget = int(float(_scalar(deck, "gpu_memory")))
    except ValueError:
        raise ValueError("gpu_memory is not a number")
    return _set_scalar(text, "gpu_memory", str(budget * domains)), False


def plan(deck, gpus, cell_size=None, regrid=False):
    """-> list of dicts {deck, inflow, angle, cases, inflow_split, grid_pinned}: at most `gpus` Cartesian sub-decks that cover every (inflow, angle) pair exactly once."""
    inflow, angle = parse_list(deck, "inflow"), parse_list(deck, "angle")
    if not inflow or not angle:
        raise ValueError("a dataset-generation deck needs inflow = [...] and angle = [...] (FX/setup.cpp:3652, 5643)")
    deck, grid_pinned = one_gpu_grid(deck, cell_size, regrid)
    gpus = max(1, int(gpus))
    angle_chunks = _chunks(angle, gpus)
    inflow_chunks = [inflow]
    if len(angle_chunks) < gpus and len(inflow) > 1:  # fewer angles than GPUs: split the inflows as well
        inflow_chunks = _chunks(inflow, gpus // len(angle_chunks))
    subs = []
    for ic in inflow_chunks:
        for ac in angle_chunks:
            text = _set_list(_set_list(deck, "inflow", ic), "angle", ac)
            subs.append(dict(deck=text, inflow=ic, angle=ac, cases=[(i, a) for i in ic for a in ac], inflow_split=len(inflow_chunks) > 1, grid_pinned=grid_pinned))
    assert len(subs) <= gpus
    return subs


def launch(deck_path, driver, gpus, devices=None, dry_run=False, env=None, cell_size=None, regrid=False, stagger_s=1.1, on_exit=None):
    """Write one sub-deck per replica next to the original deck (the project's inputs are addressed relative to it), start one driver process per GPU with
    CUDA_VISIBLE_DEVICES set, wait for all. Returns [(device, sub-deck path, return code, log path)]. `devices`: CUDA ordinals to use (default 0..gpus-1).
    Replicas get no stdin (every error path of the driver ends in wait() = std::cin.get(), FX/utilities.hpp:3197: a failing replica must exit, not block), start
    `stagger_s` apart (the driver names its console log proj_temp/<YYYYmmddHHMMSS>_lbm.log and truncates it, FX/setup.cpp:2502-2510), and are reported through
    `on_exit(device, path, return code, log)` as they finish."""
    deck_path = os.path.abspath(deck_path)
    project = os.path.dirname(deck_path)
    stem, ext = os.path.splitext(os.path.basename(deck_path))
    subs = plan(open(deck_path).read(), gpus, cell_size=cell_size, regrid=regrid)
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
        if k > 0 and stagger_s > 0:
            time.sleep(stagger_s)
        procs.append((devices[k], path, subprocess.Popen([driver, path], cwd=project, env=e, stdin=subprocess.DEVNULL, stdout=fh, stderr=subprocess.STDOUT), log, fh))
    out = {}
    while len(out) < len(procs):  # report replicas as they finish
        for k, (dev, path, p, log, fh) in enumerate(procs):
            if k in out:
                continue
            rc = None if p is None else p.poll()
            if p is not None and rc is None:
                continue
            if fh is not None:
                fh.close()
            out[k] = (dev, path, rc, log)
            if on_exit is not None and p is not None:
                on_exit(dev, path, rc, log)
        if len(out) < len(procs):
            time.sleep(0.05)
    return [out[k] for k in range(len(procs))]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("deck")
    ap.add_argument("--driver", required=True, help="the case driver binary (baseline/_ref/luw_reference_driver or a build of the reference tree against host/lbm.hpp)")
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--devices", default="", help="comma-separated CUDA ordinals (default 0..gpus-1)")
    ap.add_argument("--dry-run", action="store_true", help="write the sub-decks and print the plan, start nothing")
    ap.add_argument("--cell-size", type=float, default=None, help='pin the grid: sub-decks get mesh_control = "cell_size" with this cell size [m] (needed for gpu_memory decks written for several GPUs)')
    ap.add_argument("--regrid", action="store_true", help="gpu_memory decks written for several GPUs: scale gpu_memory by the device count instead (approximate grid)")
    a = ap.parse_args(argv)
    devices = [int(v) for v in a.devices.split(",")] if a.devices else None
    try:
        res = launch(a.deck, a.driver, a.gpus, devices, a.dry_run, cell_size=a.cell_size, regrid=a.regrid,
                     on_exit=lambda dev, path, rc, log: print(f"gpu {dev}: {os.path.basename(path)} finished, exit {rc}", flush=True))
    except ValueError as e:
        print("error:", e, file=sys.stderr)
        return 2
    for dev, path, rc, log in res:
        print(f"gpu {dev}: {os.path.basename(path)} -> {'planned' if rc is None else 'exit %d' % rc} ({log})")
    return 0 if all(rc in (None, 0) for _, _, rc, _ in res) else 1


if __name__ == "__main__":
    sys.exit(main())
