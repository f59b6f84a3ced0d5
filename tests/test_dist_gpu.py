"""The drop-in with a slab-decomposed Level 1 (gomelt_b200/dist.py): the whole driver - window moves,
single steps, subcycle blocks, dwell steps, a layer change - on 1, 2, 3 and 4 ranks reproduces the plain single-GPU run
BIT FOR BIT on every level (Level-1 temperature and state assembled from the slabs, Level 2 / 3, melt-time field).

* one rank: the same machinery (mirrors, Level-1 hook, box transfers) without a second process;
* 2 / 3 ranks sharing cuda:0 under gloo (boxes and ghost planes staged through the host): runs on a single-GPU box, so
  the N>1 protocol is checked wherever the GPU tests run;
* one rank per GPU under NCCL with the peer-memory halo exchange, when the box has >= 2 GPUs.
"""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dist_support  # noqa: E402
import driver_support  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _plain(gm, tmp, case):
    import torch

    cf = gm.computeFunctions
    cf.disable_distributed()
    os.makedirs(tmp, exist_ok=True)
    res = gm.driver.go_melt(getattr(driver_support, case)(tmp), write_final=False)
    torch.cuda.synchronize()
    L = res["Levels"]
    return {"L1T": L[1]["T0"].cpu().numpy(), "L2T": L[2]["T0"].cpu().numpy(), "L3T": L[3]["T0"].cpu().numpy(),
            "accum": res["accum_time"].cpu().numpy(), "S1": L[1]["S1"].cpu().numpy(), "counts": res["counts"]}


def _check(want, out):
    got = np.load(os.path.join(out, "owner.npz"))
    for k in ("L1T", "L2T", "L3T", "accum"):
        assert got[k].shape == want[k].shape, k
        assert np.array_equal(got[k], want[k]), (k, int((got[k] != want[k]).sum()), float(np.abs(got[k] - want[k]).max()))
    assert np.array_equal(np.load(os.path.join(out, "owner_S1.npy")), want["S1"])
    assert np.array_equal(got["counts"], np.array([want["counts"][k] for k in sorted(want["counts"])]))
    assert want["L3T"].max() > 1000.0 and int(got["solves"]) > 0 and int(got["boxes_up"]) > 0  # the run did something
    return got


@pytest.mark.parametrize("case", ["small_two_layer_input", "serpentine_input", "wide_part_input"])
def test_one_rank_distributed_equals_plain(gm, tmp_path, case):
    import torch

    cf = gm.computeFunctions
    want = _plain(gm, str(tmp_path / "plain"), case)
    try:
        cf.enable_distributed(0, 1)
        tmp = str(tmp_path / "dist")
        os.makedirs(tmp)
        res = gm.driver.go_melt(getattr(driver_support, case)(tmp), write_final=False)
        torch.cuda.synchronize()
        L = res["Levels"]
        d = cf.distOf(L)
        assert d is not None and d.is_owner and d.stats["solves"] > 0
        for k, t in (("L1T", L[1]["T0"]), ("L2T", L[2]["T0"]), ("L3T", L[3]["T0"]), ("accum", res["accum_time"]),
                     ("S1", cf.gatherL1(L, "S1"))):
            assert np.array_equal(t.cpu().numpy(), want[k]), k
    finally:
        cf.disable_distributed()


def _spawn(world, out, case, backend, one_device):
    import torch.multiprocessing as mp

    for attempt in range(3):  # (a free port can be taken between the probe and the rendezvous)
        try:
            mp.start_processes(dist_support.run_rank, args=(world, _free_port(), out, case, backend, one_device), nprocs=world,
                               join=True, start_method="spawn")
            return
        except Exception as exc:
            if "EADDRINUSE" not in str(exc) or attempt == 2:
                raise


@pytest.mark.parametrize("world,case", [(2, "small_two_layer_input"), (3, "small_two_layer_input"), (3, "wide_part_input")])
def test_ranks_sharing_one_gpu_equal_plain(gm, tmp_path, world, case):
    want = _plain(gm, str(tmp_path / "plain"), case)
    out = str(tmp_path / "dist")
    os.makedirs(out)
    _spawn(world, out, case, "gloo", True)
    got = _check(want, out)
    assert int(got["boxes_down"]) > 0


@pytest.mark.parametrize("case", ["small_two_layer_input", "wide_part_input"])
def test_one_rank_per_gpu_equals_plain(gm, tmp_path, case):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (the same protocol runs under gloo on one GPU in the test above)")
    world = min(ngpu, 4)
    want = _plain(gm, str(tmp_path / "plain"), case)
    out = str(tmp_path / "dist")
    os.makedirs(out)
    _spawn(world, out, case, "nccl", False)
    _check(want, out)


def test_output_hooks_of_a_distributed_run_see_the_assembled_level1(gm, tmp_path):
    """Records, checkpoints and the final output of a distributed run go through computeFunctions.outputView: the hooks
    get the Level-1 temperature / state / S1_storage assembled from the slabs, equal to what the plain run's hooks get."""
    import torch

    cf = gm.computeFunctions
    seen = {}

    def hooks(tag):
        seen[tag] = {"rec": [], "ckpt": [], "final": []}
        s = seen[tag]
        return {"on_record": lambda L, N, i: s["rec"].append((i, L[1]["T0"].clone(), L[1]["S1"].clone(), L[3]["T0"].clone())),
                "on_checkpoint": lambda L, a, m, t, r, N: s["ckpt"].append((t, L[1]["T0"].clone(), L[1]["S1_storage"].clone())),
                "on_final": lambda L, N: s["final"].append(L[1]["T0"].clone())}

    try:
        for tag in ("plain", "dist"):
            (cf.enable_distributed(0, 1) if tag == "dist" else cf.disable_distributed())
            d = str(tmp_path / tag)
            os.makedirs(d)
            gm.driver.go_melt(driver_support.small_two_layer_input(d), write_final=False, hooks=hooks(tag))
            torch.cuda.synchronize()
    finally:
        cf.disable_distributed()
    a, b = seen["plain"], seen["dist"]
    assert len(a["rec"]) == len(b["rec"]) > 3 and len(a["ckpt"]) == len(b["ckpt"]) >= 1 and len(a["final"]) == len(b["final"]) == 1
    for x, y in zip(a["rec"], b["rec"]):
        assert x[0] == y[0] and all(torch.equal(p, q) for p, q in zip(x[1:], y[1:]))
    for x, y in zip(a["ckpt"], b["ckpt"]):
        assert x[0] == y[0] and torch.equal(x[1], y[1]) and torch.equal(x[2], y[2])
    assert torch.equal(a["final"][0], b["final"][0])


def test_distributed_restart_from_a_checkpoint_equals_the_uninterrupted_plain_run(gm, tmp_path):
    """gm:108-129 + 390-411 with a slab-decomposed Level 1: a distributed run stopped at its first checkpoint (written in
    the single-GPU format from the assembled fields) and restarted - again distributed: every rank cuts its slab out of
    the checkpoint - ends bit for bit where the uninterrupted plain run ends."""
    import torch

    cf = gm.computeFunctions
    out = gm.output
    hooks = {k: v for k, v in out.driver_hooks().items() if k in ("on_checkpoint", "load_checkpoint", "on_layer_accum")}
    want = _plain(gm, str(tmp_path / "plain"), "small_two_layer_input")
    d = str(tmp_path / "split")
    os.makedirs(d)
    try:
        cf.enable_distributed(0, 1)
        inp = driver_support.small_two_layer_input(d)
        inp["nonmesh"]["restart_layer_num"] = 1
        first = gm.driver.go_melt(inp, hooks=hooks, write_final=False)
        assert first["stopped_at_layer_check"] and os.path.exists(os.path.join(d, "checkpoint", "Checkpoint0001", "header.json"))
        inp = driver_support.small_two_layer_input(d)
        inp["nonmesh"].update(layer_num=1, use_txt=1)
        second = gm.driver.go_melt(inp, hooks=hooks, write_final=False)
        torch.cuda.synchronize()
        L = second["Levels"]
        assert cf.distOf(L) is not None
        got = {"L1T": L[1]["T0"], "L2T": L[2]["T0"], "L3T": L[3]["T0"], "accum": second["accum_time"], "S1": cf.gatherL1(L, "S1")}
        for k, t in got.items():
            assert np.array_equal(t.cpu().numpy(), want[k]), k
    finally:
        cf.disable_distributed()
