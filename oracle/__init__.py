"""CPU oracle for the SOM-codebook hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there
only as the checker (or as the timed CPU baseline), never as the thing shipped.

Pinning status: the reference repository ships no tests, golden vectors or KATs for
this path, so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
``oracle/make_golden.py`` imports ``/root/reference/models/Codebook.py`` in the build
container, runs it on seeded inputs and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those files bit-for-bit
(same torch build) on every CPU test run.
"""
from .codebook_oracle import (  # noqa: F401
    OracleCodebook,
    patchify,
    unpatchify,
    bmu_fp64,
    distance_fp64,
    neighbourhood_two_var,
    band_half_width,
)
from .step_oracle import (  # noqa: F401
    AdamState,
    reference_step,
    closed_form_step,
    factorised_grad,
    histogram_prune,
    synthetic_fmaps,
    trained_like_codebook,
)
from .tokens_oracle import tokenize_pair_oracle  # noqa: F401
