"""CPU oracle for the STFT-family DSP hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the CPU arm that is timed beside the GPU.

PARITY UNPINNED: the reference (smdesai/mlx-swift-audio) has no golden vectors or
known-answer tests for this path, and it cannot be built or run outside macOS (Swift +
MLX + Apple frameworks), so this restatement is the only executable statement of the
reference available here.  See DESIGN.md "Oracle".
"""
