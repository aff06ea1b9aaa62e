"""CPU oracle for the ACKTR learner hot path of jrobine/actor-critic.

TEST INFRASTRUCTURE ONLY.  Nothing under ``actorcritic_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.  It restates, in NumPy / torch-CPU,
what the reference computes (citations are ``/root/reference`` file:line) so the
CUDA path can be checked against it.

Pinning status (see DESIGN.md "Oracle"):
  * preprocess / frame stack / returns / A2C loss forward: PINNED against the
    reference's own Python code run under shims (tests/golden/make_golden.py)
    and against the real ``cv2`` the reference calls.
  * gradients: pinned against torch autograd (independent derivation), not the
    reference (TensorFlow absent).
  * K-FAC arithmetic (tensorflow/kfac 0.1.x, not vendored, not installable):
    PARITY UNPINNED - restated from the published algorithm; anchored only on
    the reference's call sites (envs/atari/model.py:219-246, kfac_utils.py:38-53,
    a2c_acktr.py:243-247).
"""
