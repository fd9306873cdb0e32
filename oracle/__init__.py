"""CPU oracle for the ArcFace head and the gallery match.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the CPU arm that is timed *beside* the GPU path).  The product
package (``b200face``) never imports this module and fails loudly when its
CUDA library is missing.

Parity pinning: the reference ships no tests, so there are no reference-owned
golden vectors.  The oracle is pinned against (a) outputs of the reference's
own ``ArcMarginProduct`` / ``ArcFaceNet`` / ``compare_faces`` code executed in
the build container (``tests/golden/make_golden.py`` loads those files by path
from /root/reference and freezes inputs+outputs into ``tests/golden/*.npz``)
and (b) the one real-data fixture the reference holds,
``face_references/face_references.pkl`` (its 7x7 distance matrix is frozen in
``tests/golden/gallery_fixture.npz``).
"""
from .arcface_oracle import (  # noqa: F401
    HeadConfig,
    warmup_schedule,
    effective_margin_scale,
    l2_normalize_rows,
    arc_logits,
    smoothed_cross_entropy,
    head_forward_backward,
    hook_kappa,
    sharded_head_forward_backward,
)
from .gallery_oracle import (  # noqa: F401
    pairwise_distance_eps,
    compare_faces,
    gallery_topk,
    cosine_class_match,
    merge_topk_shards,
)
