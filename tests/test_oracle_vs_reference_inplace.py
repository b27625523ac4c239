"""Where the reference tree exists (the build container), the restatement of its CPU path (oracle/reference_path.py, what
`bench.py --impl reference` times on the GPU box) must give the reference's own answers, crop by crop, bit for bit: the
reference functions are imported in place from /root/reference (nothing is copied) and run on the same seeded crops.
Skipped where the tree is absent -- there the committed golden vectors (tests/test_oracle_golden.py) are the pin."""
import numpy as np
import pytest

from oracle import decode, ref_inplace, reference_path
from workloads import synth

pytestmark = pytest.mark.skipif(ref_inplace.modules() is None, reason="/root/reference is not available here")


@pytest.mark.parametrize("ignore_bit", [0, 2])
def test_port_equals_reference_functions(ignore_bit):
    _, _, gnd = ref_inplace.modules()
    tab, nrm, _ = synth.make_dict(16, seed=3, radius=51.0, missing_frac=0.1)
    d = decode.table_to_dict(tab, float_keys=True)
    if ignore_bit:
        d = gnd.generate_new_corres_dict(d, 16, 16 - ignore_bit)          # the reference's own ignore-bit dictionary
        port = decode.table_to_dict(decode.generate_new_corres_table(tab, 16, 16 - ignore_bit), float_keys=True)
        assert set(d) == set(port)
        for k in d:
            assert np.array_equal(np.asarray(d[k], float).ravel(), np.asarray(port[k], float).ravel(), equal_nan=True)
    assert reference_path.kind() == "reference"
    for i in range(3):
        c = synth.make_crop(tab, nrm, 1234 + i)
        lg = synth.crop_to_logits(c)
        a = reference_path.reference_pose_from_logits(lg, c["bbox"], c["K"], d, ignore_bit=ignore_bit)
        b = reference_path.reference_pose_from_logits_ref(lg, c["bbox"], c["K"], d, ignore_bit=ignore_bit)
        assert a[2] == b[2] and a[2]
        assert np.array_equal(np.asarray(a[0]), np.asarray(b[0])) and np.array_equal(np.asarray(a[1]), np.asarray(b[1]))


def test_edge_cases_match():
    tab, nrm, _ = synth.make_dict(16, seed=3, radius=51.0, missing_frac=0.1)
    d = decode.table_to_dict(tab, float_keys=True)
    c = synth.make_crop(tab, nrm, 99)
    lg = synth.crop_to_logits(c)
    empty = lg.copy(); empty[0] = -5.0                                     # no mask pixel
    few = lg.copy(); few[0] = -5.0; few[0, 10, 10:15] = 5.0                # five mask pixels: below the six the reference asks for
    for x in (empty, few):
        a = reference_path.reference_pose_from_logits(x, c["bbox"], c["K"], d)
        b = reference_path.reference_pose_from_logits_ref(x, c["bbox"], c["K"], d)
        assert a[2] is False and b[2] is False and len(a[0]) == 0 and len(b[0]) == 0
