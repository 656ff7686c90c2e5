"""CPU oracle for the sed-crnn hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the shipped package (`sed_crnn_b200/`) may import from here.  The only
legitimate importers are `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`, and there only as the checker or as the thing timed
as "the reference's CPU path" -- never as the product.

Pinning status (see DESIGN.md section "Oracle"):
  * metrics_ref      -- pinned: bit-exact against /root/reference/metrics.py executed in the
                        build container (tests/golden/metrics_kat.npz, made by
                        oracle/make_golden.py).
  * crnn_ref         -- pinned: fork-mode forward/backward/one-step outputs equal to the
                        reference's own crnn_lightning.TimePooledCRNN / sed.TimePooledCRNN /
                        FocalBCELoss run on torch-CPU fp32 (tests/golden/crnn_fork_*.npz).
  * logmel_ref       -- the reference delegates this arithmetic to librosa, which is not
                        vendored in /root/reference and not installable offline.  The
                        restatement follows librosa's published algorithm and is cross-checked
                        against two independent implementations (torchaudio mel + torch.stft),
                        but the reference ships no golden vector for it: PARITY UNPINNED
                        for the log-mel leg.
"""
