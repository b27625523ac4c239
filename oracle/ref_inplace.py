"""The reference's OWN functions of the hot path, imported IN PLACE from /root/reference (test infrastructure).

Nothing is copied: the reference tree only exists in the build container, so this is what the CPU tests there use to keep the
restatement in oracle/ pinned against the reference itself, crop by crop (tests/test_oracle_vs_reference_inplace.py), and what
`bench.py --impl reference` runs when it is launched where the reference is (`kind: "reference"`); on the GPU box the tree is
absent and the bit-identical restatement runs (`kind: "port"`).
"""
import os
import sys

ROOT = "/root/reference/zebrapose"
FILES = ["common_ops.py", "binary_code_helper/CNN_output_to_pose.py", "binary_code_helper/class_id_encoder_decoder.py",
         "binary_code_helper/generate_new_dict.py"]
_mods = None


def modules():
    """(common_ops, CNN_output_to_pose, generate_new_dict) of the reference, or None when the tree is not there.  The
    reference imports its helpers as top-level `binary_code_helper.*`, so its directory goes on sys.path (the product's
    mirrors live under zebrapose_b200.* and do not collide)."""
    global _mods
    if _mods is None:
        _mods = False
        if all(os.path.exists(os.path.join(ROOT, f)) for f in FILES):
            import importlib
            sys.path.insert(0, ROOT)
            try:
                co = importlib.import_module("common_ops")
                cnn = importlib.import_module("binary_code_helper.CNN_output_to_pose")
                gnd = importlib.import_module("binary_code_helper.generate_new_dict")
                if all(os.path.abspath(m.__file__).startswith(ROOT) for m in (co, cnn, gnd)):
                    _mods = (co, cnn, gnd)
            except Exception:
                _mods = False
            finally:
                if ROOT in sys.path:
                    sys.path.remove(ROOT)
    return _mods or None
