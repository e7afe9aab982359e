"""generative/utils/misc.py:19-26."""


def unsqueeze_right(arr, ndim: int):
    """Append 1-sized dimensions to ``arr`` up to ``ndim`` dimensions."""
    return arr[(...,) + (None,) * (ndim - arr.ndim)]


def unsqueeze_left(arr, ndim: int):
    """Prepend 1-sized dimensions to ``arr`` up to ``ndim`` dimensions."""
    return arr[(None,) * (ndim - arr.ndim)]
