"""Type placeholder: `SparseTensor` is only used in isinstance checks (Ob_propagation.py:129)."""


class SparseTensor:  # pragma: no cover
    pass
