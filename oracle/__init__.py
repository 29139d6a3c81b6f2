"""ORACLE package: CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from here.  The product package (mujoco_panda_pnp_b200) never does, and has
no CPU fallback: it raises if the CUDA library is missing.
"""
