"""TEST INFRASTRUCTURE ONLY — CPU oracle for the EE-GAN word-region attention / DAMSM path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker or the timed CPU arm.  The product (``eegan_b200``) never
imports this package and has no CPU fallback.

Parity pin: the reference (qikizh/EE-GAN) ships no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned by *executing the reference itself* in the build
container (``oracle/ref_loader.py`` imports ``/root/reference/miscc/DAMSM_losses.py``
unmodified) — see ``oracle/make_golden.py`` which wrote ``tests/golden/*.npz`` and
``tests/test_oracle_vs_reference.py`` which re-checks the oracle against the live reference
whenever ``/root/reference`` is mounted.
"""
