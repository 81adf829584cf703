"""`mqe` module paths of the reference, served by mqe_b200 (see compat/README.md)."""
