"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds the CPU checker for the gim_loftr hot path:

* `loftr_oracle.py` - a plain torch-CPU fp32 restatement of the reference algorithm
  (every function cites the reference file:line it follows);
* `ref_import.py`   - imports the UNMODIFIED reference from /root/reference (only possible in
  the build container) to pin the restatement and to generate `tests/golden/*`;
* `make_golden.py`  - the script that produced the committed golden vectors.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import anything from here.  The product (`gim_b200/`) never does.
"""
