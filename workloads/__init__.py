"""Seeded synthetic inputs of BASELINE.json's configs: dictionaries, crops, logits,
detection boxes, model vertices, pose pairs.  Input generation only -- no part of the path, no checker: bench.py, the
tests, the golden-fixture scripts and tools/ all draw their inputs from here so they see identical data."""
