"""CPU oracle for the SqueezeDet inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``squeezedet_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may.  It is a restatement, in
numpy, of the arithmetic the reference performs on its one data-parallel hot
path (TF-1.0 graph ops + numpy post-processing); every function cites the
reference file:line it follows.

PARITY PINNING:
  * numpy half (``batch_iou``, ``nms``, ``filter_prediction``, ``set_anchors``):
    PINNED — checked against the reference's own functions imported from
    ``/root/reference/src`` (``oracle.ref_import``), both live in the dev
    container and through committed fixtures in ``tests/golden/``.
  * TensorFlow half (conv / pool / batch-norm / interpret_output): **parity
    unpinned** — tensorflow 1.0 is not installable here and the reference ships
    no tests or golden vectors; the restatement follows the TF op semantics of
    SURVEY.md App. A and is cross-validated two ways (naive loops vs im2col-GEMM
    vs torch-CPU), nothing external pins it.
  * pre-processing (``oracle.preproc``: cv2 float32 INTER_LINEAR resize + mean
    subtraction, both reference orders): PINNED against the installed cv2, the
    reference's own dependency for that step (tests/test_oracle_preproc.py).
"""

from .semantics import (  # noqa: F401
    conv_geometry, conv2d, conv2d_naive, max_pool, max_pool_naive,
    batch_norm_frozen, relu,
)
from .postproc import (  # noqa: F401
    set_anchors, safe_exp, interpret_output, batch_iou, nms,
    filter_prediction,
)
from .nets import forward, NET_BUILDERS, layer_table, param_specs  # noqa: F401
