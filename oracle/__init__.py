"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference algorithm for the EDM sampling hot path (torch fp32/fp64 functional code
and NumPy), used as the parity checker by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
`--impl reference` legs.  Nothing under tqdne_b200/ may import this package: the product path is CUDA-only.

Pinning status
  * UNet / ResBlock / attention / Encoder / Decoder / EDM preconditioning / Heun sampler / MovingAverageEnvelope
    inverse / LogSpectrogram un-normalisation: PINNED against the reference's own modules executed in the
    build container (oracle/make_golden.py imports /root/reference; vectors in tests/golden/).
  * Griffin-Lim (librosa 0.11.0 `griffinlim`, third-party, absent from /root/reference and from this image):
    PARITY UNPINNED.  oracle/griffinlim_ref.py restates the published algorithm; its STFT/iSTFT are pinned
    against torch.stft/istft, the loop is not pinned against librosa itself.
  * Training step (autograd through `torch_ref.denoise` + the EDM loss), evaluation classifier embedding and Frechet
    distance: PINNED against the reference's own `LightningEDM.step` under autograd, `LithningClassifier.embed / forward`
    and `tqdne.metric.frechet_distance` (oracle/make_golden_train.py, oracle/make_golden_classifier.py).
`tools/bench_train.py --impl reference` (the CPU arm of the training measurement) is the one other executor of this package.
"""
