"""CPU oracle for the OOD-GAN-inversion hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch fp32 restatement (functional, state-dict driven) of the
reference's algorithm for the hot path named in BASELINE.json.  It exists so that the
CUDA path can be checked against something that runs anywhere.  It is NOT part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker / the CPU baseline.
The product package (``ood_gan_inversion_b200``) never imports it and fails loudly when
its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the *reference itself*,
imported in the authoring container by ``tests/golden/make_golden.py`` (committed) and
frozen under ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` replays them.

Every function cites the reference file:line (relative to /root/reference) it restates.
"""
from . import ops, stylegan, samm, ood  # noqa: F401
