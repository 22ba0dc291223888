"""CPU oracle for the AG2Vid per-frame generation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  ``ag2video_b200`` never imports this package.

The reference (roeiherz/AG2Video) is pure Python on top of PyTorch: all of the
arithmetic of the hot path lives in the third-party dependency ``torch``
(reference pin ``pytorch==1.4.0``, README.md:23; this image ships torch 2.11).
The oracle therefore restates the reference's *algorithm* (which tensors are
gathered, in which order things are summed, where masks apply) with the same
torch CPU fp32 primitives the reference calls (``F.grid_sample``,
``F.conv2d``, ``F.batch_norm``, ``F.interpolate``), plus an independent numpy
closed form for the layout support (``oracle.closed_form``).

Parity pin: the reference holds no golden vectors or tests for this path
("parity unpinned" by the reference).  The oracle is
pinned instead against outputs of the *reference code itself*, imported and run
in the build container by ``tests/golden/make_golden.py`` (script committed,
fixtures under ``tests/golden/*.pt``); ``tests/test_oracle_golden.py`` checks
every oracle function against them.
"""
