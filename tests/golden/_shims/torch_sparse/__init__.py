"""Import stub: fork model jointsrmfsparse.py:1 imports torch_sparse eagerly."""


class SparseTensor:  # never instantiated on the hot path
    pass
