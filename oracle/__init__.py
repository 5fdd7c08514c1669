"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain PyTorch fp32 / numpy) of the reference's algorithm for the hot path, used as the parity
checker.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import anything from here; the product (`rehrseg_b200/`) never does and fails loudly without its CUDA library.

Pinning status (see DESIGN.md "Oracle"):
  * In-tree reference code (models/seg_model.py, utils/seg_utils.py:176-287, utils/fba.py, utils/patch_ops.py,
    utils/rotate.py, utils/pad.py, utils/sr_utils.py:102-135, train_all.py:85-112, models/FLAVR/*): the reference
    ships NO tests, golden vectors or fixtures, so the restatements are pinned against outputs of the reference's
    own modules executed in the build container (oracle/make_golden.py -> tests/golden/*.npz, and live
    cross-checks in tests/test_oracle_vs_reference.py whenever /root/reference is present).
  * Third-party pieces the reference imports but does not vendor (dynamic_network_architectures 0.3.1 PlainConvUNet /
    UNetDecoder, nnunetv2 2.3.1 compute_gaussian and soft-dice, acvl_utils 0.2 pad_nd_image): absent from
    /root/reference and from this image -> restated from their published behaviour in oracle/third_party.py;
    PARITY UNPINNED for those pieces (nothing in the reference pins them).
"""
