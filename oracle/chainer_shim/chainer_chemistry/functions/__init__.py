from oracle.minichainer import matmul  # noqa: F401  (chainer_chemistry.functions.matmul = batched matmul)
