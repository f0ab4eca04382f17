"""Import-path shim: the reference package layout (`src.…`) re-exported from deeploopcloser_b200."""
