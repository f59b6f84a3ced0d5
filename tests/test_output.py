"""Output path: .vtr writer, reference-named save functions, raw-dump checkpoints,
toolpath seek, fused min / max monitor."""
import importlib
import os

import numpy as np
import pytest

out = importlib.import_module("gomelt_b200.output")


def _level(nodes=(5, 4, 3), seed=0):
    rng = np.random.default_rng(seed)
    nx, ny, nz = nodes
    return {"node_coords": [np.linspace(0, 1, nx, dtype=np.float32), np.linspace(0, 2, ny, dtype=np.float32),
                            np.linspace(-1, 0, nz, dtype=np.float32)],
            "nodes": list(nodes), "nn": nx * ny * nz, "h": [0.25, 2 / 3, 0.5],
            "T0": (300 + 100 * rng.random(nx * ny * nz)).astype(np.float32),
            "S1": (rng.random(nx * ny * nz) > 0.5).astype(np.float32),
            "S2": rng.random(nx * ny * nz) > 0.7, "bounds": {"ix": (np.float32(-1.0), np.float32(2.0))}}


def test_vtr_round_trip_and_layout(tmp_path):
    L = _level()
    nx, ny, nz = L["nodes"]
    p = out.write_vtr(str(tmp_path / "a"), *L["node_coords"], {"T": L["T0"], "S as cube": L["S1"].reshape(nz, ny, nx).transpose(2, 1, 0)})
    assert p.endswith(".vtr")
    head = open(p, "rb").read(400).decode(errors="replace")
    assert 'type="RectilinearGrid"' in head and f'WholeExtent="0 {nx - 1} 0 {ny - 1} 0 {nz - 1}"' in head
    (x, y, z), f = out.read_vtr(p)
    assert np.array_equal(x, L["node_coords"][0]) and np.array_equal(z, L["node_coords"][2])
    assert np.array_equal(f["T"], L["T0"])            # flat x-fastest order is the file order
    assert np.array_equal(f["S as cube"], L["S1"])    # the (nx, ny, nz) layout pyevtk takes lands in the same order
    with pytest.raises(ValueError):
        out.write_vtr(str(tmp_path / "b"), *L["node_coords"], {"T": L["T0"][:-1]})


def test_reference_save_functions_names_and_record_rule(tmp_path):
    """cF:1945-2057, 3668-3693: file names, z offsets, Level-1 record step."""
    Levels = [_level((3, 3, 3), 9), _level((5, 4, 3), 1), _level((6, 5, 4), 2), _level((7, 5, 4), 3)]
    nm = {"output_files": 1, "Level1_record_step": 3, "save_path": str(tmp_path) + "/", "layer_num": 0}
    w1 = out.saveResults(Levels, nm, 1)
    assert [os.path.basename(p) for p in w1] == ["Level1_00000001.vtr", "Level2_00000001.vtr", "Level3_00000001.vtr"]
    assert [os.path.basename(p) for p in out.saveResults(Levels, nm, 2)] == ["Level2_00000002.vtr", "Level3_00000002.vtr"]
    assert len(out.saveResults(Levels, nm, 4)) == 3       # mod(4, 3) == 1
    (x, y, z), f = out.read_vtr(w1[0])
    assert np.allclose(z, Levels[1]["node_coords"][2] - 2e-3)
    assert set(f) == {"Temperature (K)", "State (Powder/Solid)"} and np.array_equal(f["Temperature (K)"], Levels[1]["T0"])
    fin = out.saveResultsFinal(Levels, nm)
    assert [os.path.basename(p) for p in fin] == ["Level1_Final.vtr", "Level2_Final.vtr", "Level3_Final.vtr"]
    st = out.saveState(Levels[0], "Level0_", 7, nm["save_path"], 0)
    assert os.path.basename(st) == "Level0_00000007.vtr" and set(out.read_vtr(st)[1]) == {"State (Powder/Solid)"}
    nm["output_files"] = 0
    assert out.saveResults(Levels, nm, 1) == []


def test_checkpoint_round_trip_and_toolpath_seek(tmp_path):
    Levels = [_level((3, 3, 3), 9), _level((5, 4, 3), 1), _level((6, 5, 4), 2), _level((7, 5, 4), 3)]
    Levels[1]["conditions"] = {"x": [np.float32(298.15), np.float32(300.0)]}
    Levels[0]["idx"] = np.arange(12, dtype=np.int32)
    acc, mx = np.arange(27, dtype=np.float32), np.ones(27, np.float32)
    d = out.save_checkpoint(str(tmp_path / "checkpoint" / "Checkpoint0001"), Levels, acc, mx, 600, 2.5)
    assert os.path.exists(os.path.join(d, "header.json"))
    L2, acc2, mx2, time_inc, record_inc = out.load_checkpoint(d)
    assert time_inc == 600 and record_inc == 2.5 and np.array_equal(acc2, acc) and np.array_equal(mx2, mx)
    for a, b in zip(Levels, L2):
        for k in ("T0", "S1", "S2"):
            assert np.array_equal(a[k], b[k]) and a[k].dtype == b[k].dtype
        assert a["nodes"] == b["nodes"] and a["nn"] == b["nn"]
        assert all(np.array_equal(p, q) for p, q in zip(a["node_coords"], b["node_coords"]))
    assert L2[1]["conditions"]["x"][1] == np.float32(300.0) and L2[1]["conditions"]["x"][1].dtype == np.float32
    assert np.array_equal(L2[0]["idx"], Levels[0]["idx"]) and L2[0]["idx"].dtype == np.int32
    assert L2[1]["bounds"]["ix"] == (np.float32(-1.0), np.float32(2.0))
    # fixed-width toolpath rows (cP:71-74: 82 bytes per row): seek to row k of the example toolpath
    tp = os.path.join(os.path.dirname(__file__), "golden", "toolpath_example.txt")
    lines = open(tp).readlines()
    with open(tp) as fh:
        w = out.toolpath_seek(fh, 5)
        assert w == len(lines[0]) and fh.readline() == lines[5]


@pytest.mark.gpu
def test_minmax_monitor_and_device_snapshots(gm, tmp_path):
    import torch

    g = torch.Generator(device="cuda").manual_seed(0)
    x = 5000 * torch.rand(1_000_003, device="cuda", generator=g) - 1200
    o = gm.ops.minmax(x).cpu().numpy()
    assert o[0] == x.min().item() and o[1] == x.max().item() and o[2] == 0
    x[17] = float("nan"); x[999] = float("inf"); x[5] = -float("inf")
    o = gm.ops.minmax(x).cpu().numpy()
    fin = x[torch.isfinite(x)]
    assert o[0] == fin.min().item() and o[1] == fin.max().item() and o[2] == 3
    neg = -torch.rand(70000, device="cuda", generator=g) - 3
    o = gm.ops.minmax(neg).cpu().numpy()
    assert o[0] == neg.min().item() and o[1] == neg.max().item()
    # reference-named save functions on device-resident fields (pinned async snapshot)
    L = _level((33, 20, 9), 4)
    Ld = dict(L, T0=torch.as_tensor(L["T0"]).cuda(), S1=torch.as_tensor(L["S1"]).cuda())
    p = out.saveResult(Ld, "Level3_", 12, str(tmp_path) + "/", 0)
    _, f = out.read_vtr(p)
    assert np.array_equal(f["Temperature (K)"], L["T0"]) and np.array_equal(f["State (Powder/Solid)"], L["S1"])
    mm = out.level_minmax([None, Ld])
    assert mm[0][0] == L["T0"].min() and mm[0][1] == L["T0"].max() and mm[0][2] == 0
