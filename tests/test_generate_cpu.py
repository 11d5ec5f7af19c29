"""CPU: host logic of the generate-waveforms drop-in (conditioning normalisation, CSV expansion, CLI surface,
output schema) against the reference's constants and flags."""
import re
from pathlib import Path

import numpy as np
import pytest

from tqdne_b200 import generate_waveforms as gw

REF_CLI = Path("/root/reference/tqdne/generate_waveforms.py")
REF_FLAGS = ["--hypocentral_distance", "--magnitude", "--vs30", "--hypocentre_depth", "--azimuthal_gap", "--num_samples",
             "--csv", "--outfile", "--edm_checkpoint", "--autoencoder_checkpoint", "--batch_size"]


def test_cli_accepts_every_reference_flag_with_the_reference_defaults():
    p = gw.build_parser()
    a = p.parse_args(["--outfile", "o.h5"])
    assert a.batch_size == 32 and a.csv is None and a.num_samples is None and a.edm_checkpoint is None
    flags = {s for act in p._actions for s in act.option_strings}
    assert set(REF_FLAGS) <= flags
    if REF_CLI.exists():  # build container: the flag list is read from the reference itself
        ref = set(re.findall(r'"(--[a-z0-9_]+)"', REF_CLI.read_text()))
        assert ref == set(REF_FLAGS)


def test_feature_normalisation_uses_the_reference_statistics():
    c = gw.normalize_features([101.29891904350877, 142.08307872902394], [4.801697862929673] * 2, [384.7045105848187] * 2,
                              [38.359214998072] * 2, [129.92139043457396 + 89.69479051949207] * 2)
    assert c.shape == (2, 5) and c.dtype == np.float64
    assert np.allclose(c, [[0, 0, 0, 0, 1], [1, 0, 0, 0, 1]], atol=1e-12)
    if REF_CLI.exists():
        stats = np.array([[float(a), float(b)] for a, b in re.findall(r"\[\s*([0-9.]+),\s*([0-9.]+)\s*\],", REF_CLI.read_text())][:5])
        assert np.array_equal(stats, gw.SUMMARY_STATISTICS)


def test_csv_rows_repeat_num_samples_times(tmp_path):
    f = tmp_path / "grid.csv"
    f.write_text("hypocentral_distance,hypocentre_depth,magnitude,vs30,azimuthal_gap,num_samples\n"
                 "10.0,5.0,4.5,300.0,100.0,2\n50.0,8.0,6.0,700.0,130.0,3\n")
    cols = gw.read_csv_features(f)
    assert cols["magnitude"] == [4.5, 4.5, 6.0, 6.0, 6.0] and cols["vs30"][-1] == 700.0 and len(cols["azimuthal_gap"]) == 5
    with pytest.raises(ValueError, match="CSV or a full parameter set"):
        gw.generate(None, 5.0, None, None, None, 1, None, "x", 1, None, None)


def test_output_schema(tmp_path):
    feats = {"hypocentral_distance": [1.0, 2.0], "magnitude": [5.0, 5.0], "vs30": [300.0, 300.0],
             "hypocentre_depth": [10.0, 10.0], "azimuthal_gap": [130.0, 130.0]}
    w = np.random.default_rng(0).standard_normal((2, 3, 4064))
    out = gw.write_outputs(tmp_path / "w.h5", feats, w)
    assert out.endswith("w.h5") and open(out, "rb").read(8) == b"\x89HDF\r\n\x1a\n"
    z = gw.read_outputs(out)
    assert set(z) == {"hypocentral_distance", "magnitude", "vs30s", "hypocentre_depth", "azimuthal_gap", "waveforms"}
    assert z["waveforms"].shape == (2, 3, 4064) and z["waveforms"].dtype == np.float32
    assert np.array_equal(z["waveforms"], w.astype(np.float32))
    assert z["magnitude"].dtype == np.float64 and list(z["vs30s"]) == [300.0, 300.0]


def test_minimal_hdf5_writer_structures(tmp_path):
    """hdf5_min writes the version-0 superblock / version-1 group structures of the HDF5 file format specification
    (no libhdf5 in this image to open the file with: the reader below follows the addresses in the file)."""
    import struct

    from tqdne_b200 import hdf5_min

    rng = np.random.default_rng(1)
    data = {f"d{i:02d}": rng.standard_normal((i + 1, 3)).astype(np.float32 if i % 2 else np.float64) for i in range(11)}
    data["counts"] = np.arange(7, dtype=np.int64)
    data["empty_tail"] = np.zeros((0, 4), np.float32)
    data["bytes"] = np.arange(5, dtype=np.uint8)
    path = tmp_path / "m.h5"
    hdf5_min.write(path, data)           # 14 datasets: two symbol-table nodes under the group B-tree
    raw = open(path, "rb").read()
    assert raw[:8] == hdf5_min.SIGNATURE and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    leaf_k, internal_k = struct.unpack_from("<HH", raw, 16)
    assert (leaf_k, internal_k) == (4, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and free == hdf5_min.UNDEF and drv == hdf5_min.UNDEF and eof == len(raw)
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert raw[btree:btree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP"
    assert struct.unpack_from("<H", raw, btree + 6)[0] == 2      # two children
    back = hdf5_min.read(path)
    assert list(back) == sorted(data)                              # symbol-table order = byte order of the names
    for k, v in data.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v)
    with pytest.raises(TypeError):
        hdf5_min.write(tmp_path / "bad.h5", {"c": np.zeros(2, np.complex64)})
    with pytest.raises(ValueError):
        hdf5_min.write(tmp_path / "bad.h5", {"a/b": np.zeros(2)})


def test_minimal_hdf5_file_opens_with_h5py_when_available(tmp_path):
    """wherever h5py (libhdf5) IS installed, a file written by hdf5_min must open with it and give back the
    same datasets.  Skipped in this image (no libhdf5); runs on any machine that has h5py."""
    h5py = pytest.importorskip("h5py")
    from tqdne_b200 import hdf5_min

    rng = np.random.default_rng(2)
    data = {"waveforms": rng.standard_normal((5, 3, 64)).astype(np.float32), "magnitude": rng.standard_normal(5),
            "vs30s": np.arange(5, dtype=np.float64), "counts": np.arange(9, dtype=np.int64)}
    path = tmp_path / "w.h5"
    hdf5_min.write(path, data)
    with h5py.File(path, "r") as f:
        assert sorted(f.keys()) == sorted(data)
        for k, v in data.items():
            assert f[k].dtype == v.dtype and f[k].shape == v.shape and np.array_equal(f[k][...], v)


def test_reference_checkpoint_loads_without_the_reference_package(monkeypatch):
    """tests/golden/tiny_edm_reference.ckpt was written with the reference's own classes (oracle/make_golden_ckpt.py):
    its hyper_parameters pickle `tqdne.edm.EDM`.  The loader must resolve it to tqdne_b200.edm.EDM with no `tqdne`
    package importable (SURVEY 8(b) weights / on-disk row, 8(f) rank 3)."""
    import sys
    from pathlib import Path

    import torch

    import tqdne_b200 as tq
    from tqdne_b200.lightning_shim import read_checkpoint

    for k in [k for k in sys.modules if k == "tqdne" or k.startswith("tqdne.")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.setattr(sys, "path", [p for p in sys.path if "reference" not in p])
    path = Path(__file__).parent / "golden" / "tiny_edm_reference.ckpt"
    raw = read_checkpoint(path)
    assert type(raw["hyper_parameters"]["edm"]).__module__ == "tqdne_b200.edm"
    edm = tq.LightningEDM.load_from_checkpoint(path)
    assert edm.num_sampling_steps == 7 and edm.deterministic_sampling is True
    assert edm.edm.sigma_min == 0.004 and edm.edm.sigma_max == 40.0 and edm.edm.rho == 7.0
    sd = edm.state_dict()
    assert set(sd) == set(raw["state_dict"])
    assert all(torch.equal(sd[k], raw["state_dict"][k]) for k in sd)
    sig = edm.edm.sampling_sigmas(7)
    assert abs(float(sig[0]) - 40.0) < 1e-4 and abs(float(sig[-2]) - 0.004) < 1e-6 and float(sig[-1]) == 0.0
    ema = tq.LightningEDM.load_from_checkpoint(path, use_ema=True)
    k = "unet.time_mlp.0.weight"
    assert torch.equal(ema.state_dict()[k], raw["ema_state"][k]) and not torch.equal(ema.state_dict()[k], sd[k])


def test_bench_clock_sampler_filters_to_the_timed_window():
    """bench.ClockSampler (host logic): samples outside the timed window are ignored, throttle reasons are collected, and an
    unparsable timestamp falls back to all samples instead of leaving the JSON line without clocks."""
    import datetime
    import time

    import bench

    now = time.time()
    fmt = lambda t: datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]  # noqa: E731

    class FakeProc:
        def __init__(self, text):
            self.text = text

        def terminate(self):
            pass

        def communicate(self, timeout=None):
            return self.text, ""

    rows = [f"{fmt(now - 30)}, 345, 1965, 120.0, 0x0, Not Active, Not Active, Not Active, Not Active",
            f"{fmt(now - 1.0)}, 1700, 1965, 950.0, 0x4, Not Active, Not Active, Not Active, Active",
            f"{fmt(now - 0.6)}, 1690, 1965, 970.0, 0x4, Not Active, Not Active, Not Active, Active",
            f"{fmt(now - 0.2)}, 1710, 1965, 960.0, 0x4, Not Active, Not Active, Not Active, Active"]
    c = bench.ClockSampler(0)
    c.proc = FakeProc("\n".join(rows))
    got = c.stop(now - 2, now)
    assert got["sm_mhz"] == 1700.0 and got["samples"] == 3 and got["reasons"] == ["sw_power_cap"] and got["power_w_max"] == 970.0
    c.proc = FakeProc("\n".join(r.replace("/", "-") for r in rows))      # timestamp format nvidia-smi does not use
    assert c.stop(now - 2, now)["samples"] == 4
    c.proc = FakeProc("")
    assert c.stop(now - 2, now)["reasons"] == ["no samples"]
