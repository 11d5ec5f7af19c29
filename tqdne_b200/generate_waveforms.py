"""`generate-waveforms` on the B200 engine -- drop-in for tqdne/generate_waveforms.py (reference :67-268).

Same command line (`--hypocentral_distance --magnitude --vs30 --hypocentre_depth --azimuthal_gap --num_samples
--csv --outfile --edm_checkpoint --autoencoder_checkpoint --batch_size`), same conditioning normalisation
(hard-coded data-set statistics, reference :128-159, feature order dist, mag, vs30, depth, gap), same batch loop
`edm.sample([b, 3, 128, 128], cond) -> LogSpectrogram.invert_representation` (:186-193) and the same output
datasets (`hypocentral_distance, magnitude, vs30s, hypocentre_depth, azimuthal_gap, waveforms[N, 3, 4064]`, :177-184).

Differences, all explicit:
  * the whole batch loop runs on the GPU (sampler, decoder AND Griffin-Lim); under `torchrun` the rows shard
    contiguously over the ranks and rank 0 gathers and writes (no collective on the data path);
  * no network: checkpoints are never downloaded.  Without checkpoints `--random_init` builds the named
    architecture with seeded random weights (useful for benchmarking only);
  * the HDF5 file is written with h5py when it is installed, otherwise by the built-in minimal HDF5 writer
    (`tqdne_b200/hdf5_min.py`: same datasets, dtypes and names, readable by any HDF5 library).
Engine extras: `--precision {bf16,fp32}`, `--seed`, `--num_sampling_steps`, `--random_init`.
"""

from __future__ import annotations

import argparse
import csv as _csv
import os
from pathlib import Path

import numpy as np
import torch

from . import sharding
from .architectures import get_2d_autoencoder_configs, get_2d_unet_config
from .autoencoder import LightningAutoencoder
from .config import LatentSpectrogramConfig
from .edm import LightningEDM
from .utils import get_device

# summary statistics (mean, std) of the conditioning features of the reference's training set
# (reference generate_waveforms.py:128-136); order: hypocentral_distance, magnitude, vs30, hypocentre_depth, azimuthal_gap
SUMMARY_STATISTICS = np.array(
    [
        [101.29891904350877, 40.78415968551517],
        [4.801697862929673, 0.7146698731358634],
        [384.7045105848187, 220.11269086015872],
        [38.359214998072, 22.472499592355014],
        [129.92139043457396, 89.69479051949207],
    ]
)
CSV_COLUMNS = ("hypocentral_distance", "magnitude", "vs30", "hypocentre_depth", "azimuthal_gap", "num_samples")


def read_csv_features(path) -> dict:
    """Rows `hypocentral_distance,magnitude,vs30,hypocentre_depth,azimuthal_gap,num_samples`, each repeated
    `num_samples` times (reference :82-91: df.loc[df.index.repeat(df.num_samples)])."""
    cols = {k: [] for k in CSV_COLUMNS[:-1]}
    with open(path, newline="") as f:
        for row in _csv.DictReader(f):
            reps = int(float(row["num_samples"]))
            for k in cols:
                cols[k] += [float(row[k])] * reps
    return cols


def normalize_features(hypocentral_distances, magnitudes, vs30s, hypocentre_depths, azimuthal_gaps) -> np.ndarray:
    """z-score the five features with the data-set statistics -> cond [N, 5] float64 (reference :138-159)."""
    feats = [hypocentral_distances, magnitudes, vs30s, hypocentre_depths, azimuthal_gaps]
    return np.stack([(np.array(f) - SUMMARY_STATISTICS[i, 0]) / SUMMARY_STATISTICS[i, 1] for i, f in enumerate(feats)], axis=1)


def load_models(edm_checkpoint, autoencoder_checkpoint, device, random_init=False, seed=0, num_sampling_steps=None):
    if edm_checkpoint is None and autoencoder_checkpoint is None:
        if not random_init:
            raise ValueError("no checkpoints given and this build cannot download them (no network): pass "
                             "--edm_checkpoint/--autoencoder_checkpoint, or --random_init for seeded random weights")
        config = LatentSpectrogramConfig()
        torch.manual_seed(seed)
        enc_cfg, dec_cfg = get_2d_autoencoder_configs(config)
        ae = LightningAutoencoder(enc_cfg, dec_cfg, {})
        edm = LightningEDM(get_2d_unet_config(config, config.latent_channels, config.latent_channels), {}, autoencoder=ae)
        # zero_module initialises the last conv of every block to 0: give them seeded values too
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for prm in edm.parameters():
                if prm.numel() > 1 and not bool(prm.any()):
                    prm.copy_(torch.randn(prm.shape, generator=g) * (0.5 / max(1.0, float(prm[0].numel())) ** 0.5))
            # keep the random decoder's output inside the normalised log-spectrogram range [-1, 1] (exp() of anything
            # larger overflows the waveform scale)
            out = ae.decoder.output_layer
            out.weight.mul_(0.05)
            out.bias.mul_(0.05)
    elif edm_checkpoint is None or autoencoder_checkpoint is None:
        raise ValueError("Either both or none of the checkpoints must be provided.")
    else:
        ae = LightningAutoencoder.load_from_checkpoint(Path(autoencoder_checkpoint))
        edm = LightningEDM.load_from_checkpoint(Path(edm_checkpoint), autoencoder=ae)
    if num_sampling_steps is not None:
        edm.num_sampling_steps = num_sampling_steps
    return edm.to(device).eval()


def write_outputs(outfile, features: dict, waveforms: np.ndarray) -> str:
    """HDF5 with the reference's dataset names and dtypes (:177-184; note the key "vs30s"): 1-D float64 features,
    `waveforms` float32 -- through h5py when installed, else through the built-in writer (hdf5_min)."""
    names = {"hypocentral_distance": "hypocentral_distance", "magnitude": "magnitude", "vs30": "vs30s",
             "hypocentre_depth": "hypocentre_depth", "azimuthal_gap": "azimuthal_gap"}
    data = {names[k]: np.array(v) for k, v in features.items()}
    data["waveforms"] = np.asarray(waveforms, dtype=np.float32)
    try:
        import h5py
    except ImportError:
        from . import hdf5_min

        hdf5_min.write(outfile, data)
        return str(outfile)
    with h5py.File(outfile, "w") as f:
        for k, v in data.items():
            f.create_dataset(k, data=v)
    return str(outfile)


def read_outputs(path) -> dict:
    """{dataset name: array} of a generate-waveforms output file (h5py when installed, else hdf5_min.read)."""
    try:
        import h5py
    except ImportError:
        from . import hdf5_min

        return hdf5_min.read(path)
    with h5py.File(path) as f:
        return {k: f[k][:] for k in f.keys()}


@torch.no_grad()
def generate(hypocentral_distance, magnitude, vs30, hypocentre_depth, azimuthal_gap, num_samples, csv, outfile, batch_size,
             edm_checkpoint, autoencoder_checkpoint, *, precision="bf16", seed=0, random_init=False, num_sampling_steps=None,
             edm=None):
    if csv:
        print("using csv data")
        features = read_csv_features(csv)
    elif all(c is not None for c in [hypocentral_distance, magnitude, vs30, hypocentre_depth, azimuthal_gap, num_samples]):
        print("using command line input data")
        features = {"hypocentral_distance": [hypocentral_distance] * num_samples, "magnitude": [magnitude] * num_samples,
                    "vs30": [vs30] * num_samples, "hypocentre_depth": [hypocentre_depth] * num_samples,
                    "azimuthal_gap": [azimuthal_gap] * num_samples}
    else:
        raise ValueError("provide either a CSV or a full parameter set")
    cond = normalize_features(features["hypocentral_distance"], features["magnitude"], features["vs30"],
                              features["hypocentre_depth"], features["azimuthal_gap"])
    n_total = len(cond)

    rank, world = sharding.world()
    device = torch.device(get_device(), int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(device)
    config = LatentSpectrogramConfig()
    if edm is None:
        print("loading models...")
        edm = load_models(edm_checkpoint, autoencoder_checkpoint, device, random_init, seed, num_sampling_steps)
    edm.set_engine_precision(precision)

    lo, hi = sharding.shard_bounds(n_total, rank, world)
    print(f"generating waveforms {lo}..{hi} of {n_total} using {device}...")
    # The shard is processed in rounds of `batch_size` rows per rank; every round is gathered to rank 0 (the one
    # collective of the path) and copied into the host array right away, so device memory holds one batch per rank,
    # not the whole job, and the reference's streaming behaviour (generate_waveforms.py:186-193 writes batch by batch)
    # is kept.  Waveforms leave the fp64 Griffin-Lim kernel and are stored as float32 like the reference's dataset.
    shards = [sharding.shard_bounds(n_total, r, world) for r in range(world)]
    rounds = max(-(-(h - l) // batch_size) for l, h in shards) if n_total else 0
    host = np.empty((n_total, 3, config.t), dtype=np.float32) if rank == 0 else None
    for k in range(rounds):
        i, j = min(hi, lo + k * batch_size), min(hi, lo + (k + 1) * batch_size)
        if j > i:
            c = torch.tensor(cond[i:j], device=device, dtype=torch.float32)
            noise = sharding.global_noise(edm.latent_shape((j - i, 3, 128, 128))[1:], i, j, seed, device)
            sample = edm.sample([j - i, 3, 128, 128], cond=c, noise=noise)
            wav = config.representation.invert_representation_device(sample).to(torch.float32)
        else:
            wav = torch.empty((0, 3, config.t), device=device, dtype=torch.float32)
        spans = [(min(h, l + k * batch_size), min(h, l + (k + 1) * batch_size)) for l, h in shards]
        parts = sharding.gather_ragged(wav, [b - a for a, b in spans])
        if rank == 0:
            for (a, b), part in zip(spans, parts):
                if b > a:
                    host[a:b] = sharding.to_host(part).numpy()
    if rank == 0:
        out = write_outputs(outfile, features, host)
        print(f"done! -> {out}")
        return out
    return None


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Generate waveforms using the trained EDM model (B200 engine).",
                                     formatter_class=argparse.RawTextHelpFormatter)
    parser.add_argument("--hypocentral_distance", type=float, default=None, help="hypocentral distance in km")
    parser.add_argument("--magnitude", type=float, default=None, help="magnitude of the earthquake")
    parser.add_argument("--vs30", type=float, default=None,
                        help="average shear-wave velocity in the top 30 m of the site in m/s")
    parser.add_argument("--hypocentre_depth", type=float, default=None, help="hypocentre depth in km")
    parser.add_argument("--azimuthal_gap", type=float, default=None, help="azimuthal gap in degrees")
    parser.add_argument("--num_samples", type=int, default=None, help="number of samples to generate")
    parser.add_argument("--csv", type=str, default=None, help="csv file with args")
    parser.add_argument("--outfile", type=str, required=True, help="Output file name with generated waveforms")
    parser.add_argument("--edm_checkpoint", type=str, required=False, help="EDM checkpoint (Lightning .ckpt)")
    parser.add_argument("--autoencoder_checkpoint", type=str, required=False, help="Autoencoder checkpoint (Lightning .ckpt)")
    parser.add_argument("--batch_size", type=int, default=32, help="Batch size per device.")
    # engine extras
    parser.add_argument("--precision", default="bf16", choices=["bf16", "fp32"], help="tensor-core bf16 or FFMA fp32 parity mode")
    parser.add_argument("--seed", type=int, default=0, help="noise seed (noise is a function of the global sample index)")
    parser.add_argument("--num_sampling_steps", type=int, default=None, help="override the checkpoint's Heun step count")
    parser.add_argument("--random_init", action="store_true", help="seeded random weights when no checkpoints are given")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    import torch.distributed as dist

    if int(os.environ.get("WORLD_SIZE", 1)) > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    try:
        generate(args.hypocentral_distance, args.magnitude, args.vs30, args.hypocentre_depth, args.azimuthal_gap,
                 args.num_samples, args.csv, args.outfile, args.batch_size, args.edm_checkpoint, args.autoencoder_checkpoint,
                 precision=args.precision, seed=args.seed, random_init=args.random_init,
                 num_sampling_steps=args.num_sampling_steps)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
