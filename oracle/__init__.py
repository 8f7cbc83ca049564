"""CPU oracle for the Go2 convex-MPC hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in the shipped package may import this directory.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker / the CPU arm.

Parity status: the reference (ltinphan/convex-mpc-unitree-go2 v0.3.0) ships no tests,
fixtures or golden vectors, and its solver stack (CasADi 3.6.7 -> OSQP) is not installable
here.  The pure NumPy/SciPy pieces of the reference (``gait.py:26-37``,
``com_trajectory.py:15-25,213-286``) ARE executed in the authoring container through a stub
``go2_robot_data`` module and frozen into ``tests/golden/`` (``tests/golden/make_golden.py``),
so rows a1-a4 of SURVEY.md section 8 are pinned against reference outputs.  The QP solve itself
(row a13, CasADi->OSQP) is **parity unpinned**: it is a restatement of OSQP's published
algorithm (Stellato et al. 2020), certified by independent KKT checks, not by OSQP outputs.
"""
