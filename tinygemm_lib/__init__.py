"""Import-compatible alias of the reference's `tinygemm_lib` package (functional.py, utils.py),
backed by any4_b200."""
