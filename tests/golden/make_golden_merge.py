"""Golden fixture for tools_for_BOP/merge_csv.py, produced by the REFERENCE's own `main` (imported unmodified from
/root/reference/zebrapose, build container only) on three result files written by the reference's own write_cvs.
The reference concatenates in directory-scan order; `glob.glob` is wrapped to return sorted paths for the duration of
the call so that the fixture is reproducible (the mirror sorts as well).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_merge.py
"""
import glob
import os
import sys
import tempfile

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/zebrapose")

from workloads import synth_eval


def write_inputs(root, write_cvs):
    """three per-object files under <root>/<dataset>/<object>/; file 2 has integer scores only (dtype inference)"""
    for k, (ds, obj, obj_id, seed) in enumerate(synth_eval.MERGE_FILES):
        d = os.path.join(root, ds, obj)
        os.makedirs(d, exist_ok=True)
        scene, img, Rs, ts, scores = synth_eval.make_csv_rows(seed)
        if k == 1:
            scores = [1, 1, 1, -1, 1, 1, 1]
        write_cvs(d, "%s_%s" % (ds, obj), obj_id, scene, img, Rs, ts, scores)


def main():
    from tools_for_BOP import write_to_cvs, merge_csv
    with tempfile.TemporaryDirectory() as tmp:
        root = os.path.join(tmp, "results") + os.sep
        write_inputs(root, write_to_cvs.write_cvs)
        real = glob.glob
        merge_csv.glob.glob = lambda *a, **k: sorted(real(*a, **k))
        try:
            merge_csv.main(root, os.path.join(tmp, "merged.csv"))
        finally:
            merge_csv.glob.glob = real
        data = open(os.path.join(tmp, "merged.csv"), "rb").read()
    open(os.path.join(HERE, "golden_merged_v1.csv"), "wb").write(data)
    print(len(data), "bytes;", data[:80])


if __name__ == "__main__":
    main()
