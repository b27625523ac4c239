"""Drop-in for the reference package `binary_code_helper` (zebrapose/binary_code_helper/): same module and function
names, same signatures, B200 device path underneath.  Put /root/repo/zebrapose_b200 (and /root/repo) on sys.path
ahead of the reference's zebrapose/ directory and test.py / test_vivo.py import this instead."""
