"""CPU oracle for the LInKs lifting hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch/NumPy *restatement* of the reference algorithm
(Aswarin/LInKs-3D-Human-Pose-Estimation) for the path SURVEY.md section 8 scopes.
Every function cites the reference file:line it follows.

Rules (tier contract, item 3):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import this package -- as the checker / the timed
    CPU arm, never as the product.  The product (``links_b200``) never imports it and
    fails loudly when the CUDA library is missing.
  * Pinning: the reference ships **no tests, golden vectors or fixtures** (SURVEY 4).
    ``oracle/gen_golden.py`` imports the reference's own ``utils/*.py`` from
    ``/root/reference`` (build container only) and (i) asserts the restatement equals
    it on seeded inputs, (ii) freezes reference outputs into ``tests/golden/*.npz``.
    So lifters/occluders, index maps, rotation, projection, bone lengths and both
    metrics files are pinned **against outputs of the reference itself**.
  * The normalising flow lives in third-party FrEIA (github.com/VLL-HD/FrEIA,
    ``FrEIA.framework.SequenceINN`` / ``FrEIA.modules.AllInOneBlock``; version not
    pinned by the reference, not installed, no network).  ``oracle/flow.py`` restates
    its published algorithm; for that component **parity is unpinned** -- only
    self-consistency known-answer tests (invertibility, log-det == slogdet of the
    autograd Jacobian, identity-at-init) anchor it.
  * The step bodies live inside non-importable scripts (module-level argparse /
    wandb / torch.load); ``oracle/steps.py`` restates them line range by line range
    on top of the pinned pieces -- and is itself pinned against the reference's OWN
    step code: ``gen_golden.py::reference_training_steps`` cuts ``training_step`` /
    ``validation_step`` (and ``combine_pose_and_limb``) out of each script's AST and
    executes them in the build container on a stand-in ``self`` that carries the
    reference's networks, the only substitution being the oracle flow in place of the
    absent FrEIA modules (``tests/golden/ref_steps.npz``,
    ``tests/test_oracle_steps_pinned.py``: losses to 2e-5, gradients to 2e-4).
"""
