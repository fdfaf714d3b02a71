"""CPU oracle for the affinity-prediction hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, what the reference computes on the path
``inference.predict(img, model, affinity_mode=True, ...)``
(reference src/aind_exaspim_neuron_segmentation/inference.py:29-126 and
machine_learning/unet3d.py:16-336).  It exists so that the CUDA path can be
checked against an independent implementation.

Rules (enforced by tests/test_layout.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * nothing under ``aind_exaspim_neuron_segmentation_b200/`` imports it -- the
    product path has no CPU fallback and fails loudly without its CUDA library.

Pinning: the reference ships no golden vectors (tests/test_example.py:6-12 is a
placeholder).  The oracle is therefore pinned against outputs of the reference
itself, imported unmodified in the build container by
``tests/golden/make_golden.py`` (stub modules for its un-installed I/O
dependencies); the resulting vectors are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` replays them without the reference.
"""
