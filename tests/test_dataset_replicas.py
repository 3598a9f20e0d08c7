"""CPU: dataset generation as one-process-per-GPU replicas (FX/setup.cpp:5690-5753): splitting a .luwdg deck's inflow x angle list into Cartesian
sub-decks and pinning one case driver process to each GPU. A stand-in driver script records what each replica was given."""
import os
import stat
import sys

import pytest

from latticeurbanwind_b200 import dataset_replicas as R

DECK = """// LUW deck
casename = DLUTcase
// CFD Controls
n_gpu = [2, 1, 1]
mesh_control = "gpu_memory"
gpu_memory = 40000
// Batch
inflow = [2.5, 5, 7.5, 10]
angle = [0, 22.5, 45, 67.5, 90, 112.5, 135, 157.5, 180, 202.5, 225, 247.5, 270, 292.5, 315, 337.5]
run_nstep = 100
"""


CS = dict(cell_size=4.0)  # the deck above is a gpu_memory deck for two devices: the grid has to be pinned (see test_gpu_memory_decks_need_a_pinned_grid)


def _cover(subs):
    return sorted(c for s in subs for c in s["cases"])


@pytest.mark.parametrize("gpus", [1, 2, 3, 8, 16, 64])
def test_every_case_exactly_once(gpus):
    subs = R.plan(DECK, gpus, **CS)
    want = sorted((i, a) for i in R.parse_list(DECK, "inflow") for a in R.parse_list(DECK, "angle"))
    assert len(want) == 64 and _cover(subs) == want
    assert 1 <= len(subs) <= gpus
    sizes = [len(s["cases"]) for s in subs]
    assert max(sizes) - min(sizes) <= len(R.parse_list(DECK, "inflow"))  # balanced up to one angle (x all inflows)
    for s in subs:
        assert "n_gpu = [1, 1, 1]" in s["deck"] and "casename = DLUTcase" in s["deck"] and "run_nstep = 100" in s["deck"]
        assert R.parse_list(s["deck"], "inflow") == s["inflow"] and R.parse_list(s["deck"], "angle") == s["angle"]


def test_angles_are_split_before_inflows():
    """Up to 16 GPUs every replica keeps the complete inflow list: max(inflow), which the driver reads before its case loop (FX/setup.cpp:3658), is unchanged."""
    for gpus in (2, 8, 16):
        for s in R.plan(DECK, gpus, **CS):
            assert s["inflow"] == ["2.5", "5", "7.5", "10"] and not s["inflow_split"]
    subs = R.plan(DECK, 64, **CS)
    assert len(subs) == 64 and all(s["inflow_split"] and len(s["cases"]) == 1 for s in subs)


def test_gpu_memory_decks_need_a_pinned_grid():
    """mesh_control = "gpu_memory" fits the cell size to the PER-DEVICE budget (FX/setup.cpp:371-405): n_gpu = [2, 1, 1] -> [1, 1, 1] alone would coarsen the grid."""
    with pytest.raises(ValueError, match="gpu_memory"):
        R.plan(DECK, 8)
    for s in R.plan(DECK, 8, cell_size=4.0):
        assert 'mesh_control = "cell_size"' in s["deck"] and "cell_size = 4.0" in s["deck"] and s["grid_pinned"] and "n_gpu = [1, 1, 1]" in s["deck"]
    for s in R.plan(DECK, 8, regrid=True):
        assert "gpu_memory = 80000" in s["deck"] and 'mesh_control = "gpu_memory"' in s["deck"] and not s["grid_pinned"]
    single = DECK.replace("n_gpu = [2, 1, 1]", "n_gpu = [1, 1, 1]")
    assert all(s["grid_pinned"] and "gpu_memory = 40000" in s["deck"] for s in R.plan(single, 4))  # written for one device: nothing to pin
    fixed = DECK.replace('mesh_control = "gpu_memory"', 'mesh_control = "cell_size"\ncell_size = 5.0')
    assert all(s["grid_pinned"] and "cell_size = 5.0" in s["deck"] for s in R.plan(fixed, 4))


def test_literals_and_single_case_decks_survive():
    one = "inflow = [5]\nangle = [270]\nn_gpu = [2, 1, 1]\n"
    subs = R.plan(one, 8)
    assert len(subs) == 1 and subs[0]["cases"] == [("5", "270")]
    with pytest.raises(ValueError):
        R.plan("angle = [0]\n", 2)


def test_launch_pins_one_process_per_gpu(tmp_path):
    deck = tmp_path / "conf.luwdg"
    deck.write_text(DECK)
    driver = tmp_path / "fake_driver.py"
    driver.write_text("#!%s\nimport os, sys\nprint('GPU', os.environ.get('CUDA_VISIBLE_DEVICES'), 'DECK', sys.argv[1], 'CWD', os.getcwd())\nprint(open(sys.argv[1]).read())\n" % sys.executable)
    driver.chmod(driver.stat().st_mode | stat.S_IXUSR)
    seen_exit = []
    res = R.launch(str(deck), str(driver), 4, devices=[4, 5, 6, 7], stagger_s=0.0, on_exit=lambda dev, path, rc, log: seen_exit.append((dev, rc)), **CS)
    assert sorted(seen_exit) == [(4, 0), (5, 0), (6, 0), (7, 0)]
    assert [d for d, _, _, _ in res] == [4, 5, 6, 7] and all(rc == 0 for _, _, rc, _ in res)
    seen = []
    for dev, path, rc, log in res:
        text = open(log).read()
        assert f"GPU {dev} DECK {path} CWD {tmp_path}" in text.splitlines()[0]
        seen += [(i, a) for i in R.parse_list(text, "inflow") for a in R.parse_list(text, "angle")]
    assert sorted(seen) == sorted((i, a) for i in R.parse_list(DECK, "inflow") for a in R.parse_list(DECK, "angle"))
    assert R.main([str(deck), "--driver", str(driver), "--gpus", "2", "--dry-run", "--cell-size", "4.0"]) == 0
    assert R.main([str(deck), "--driver", str(driver), "--gpus", "2", "--dry-run"]) == 2  # refuses to coarsen the grid silently
