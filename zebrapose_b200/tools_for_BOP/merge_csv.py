"""Drop-in for zebrapose/tools_for_BOP/merge_csv.py:6-13: one BOP submission file out of the per-object result files
that write_to_cvs.write_cvs left under `<input_dir><dataset>/<object>/*.csv`.

Same call (`main(input_dir, output_fn)`, `--input_dir/--output_fn` on the command line), same file: the reference
round-trips every file through pandas (`read_csv` -> `concat` -> `to_csv(index=False, encoding='utf-8-sig')`), which
is what fixes the bytes -- a UTF-8 BOM, per-file dtype inference of the `score` column (an all-integer file
turns into floats as soon as one other file holds a fractional score) and pandas' float repr -- so the mirror uses the same
pandas calls rather than re-deriving them.  One deliberate difference: the reference concatenates in `glob` (directory
scan) order, which is not defined; here the files are taken in sorted path order so the output is reproducible."""
import argparse
import glob

import pandas as pd


def result_files(input_dir):
    """the reference's pattern: input_dir is used as a string prefix (pass it with its trailing separator)"""
    return sorted(glob.glob(input_dir + "*/*/*.csv"))


def main(input_dir, output_fn):
    files = result_files(input_dir)
    print(files)
    if not files:
        raise ValueError("No objects to concatenate")          # what pandas.concat raises in the reference
    merged = pd.concat([pd.read_csv(f) for f in files])
    merged.to_csv(output_fn, index=False, encoding="utf-8-sig")
    return len(files)


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="merge per-object BOP csv files")
    ap.add_argument("--input_dir", type=str)
    ap.add_argument("--output_fn", type=str)
    a = ap.parse_args()
    main(a.input_dir, a.output_fn)
