from fastfp_b200.fastfp import FastFp  # noqa: F401
