"""CPU oracle for the detection hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch fp32 ops, same operation order) of the
reference's hot path: ``Darknet.forward`` (src/darknet.py:199-253),
``predict_transform`` (src/util.py:175-239), ``write_results`` (src/util.py:242-346),
``bbox_iou`` (src/util.py:120-153) and ``confidence_mask`` (src/util.py:106-117).

It exists to *check* the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing in
``realtimeobjectdetection_b200`` imports it; the product path has no CPU fallback and
fails loudly when ``librtod.so`` is missing.

Parity pinning: the restatement is compared ``torch.equal`` against the *unmodified
reference imported from /root/reference* by ``tests/golden/make_golden.py`` (run in the
build container, where the reference is mounted); the inputs and reference outputs
of that run are committed under ``tests/golden/*.npz`` so that the pinning travels to
the GPU box, where /root/reference does not exist.  The reference itself ships no
tests; its only golden artefact, ``det/metrics.json``, is used as a row-format and
NMS-idempotence known-answer test (tests/test_oracle_golden.py).
"""
from .detect_port import (bbox_iou, confidence_mask, predict_transform,  # noqa: F401
                          write_results)
from .darknet_port import DarknetPort, parse_cfg  # noqa: F401
from . import post_port, prep_port  # noqa: F401
