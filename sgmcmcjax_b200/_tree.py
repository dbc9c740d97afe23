"""Minimal pytree flatten / unflatten with jax's leaf order (lists and tuples in order,
dict keys sorted) - the reference uses jax.tree_util for this
(sgmcmcjax/diffusion_util.py:48-52)."""
from __future__ import annotations


def flatten(tree):
    """-> (leaves, treedef); treedef is a nested description usable by unflatten."""
    leaves = []

    def rec(node):
        if isinstance(node, (list, tuple)):
            return (type(node), [rec(c) for c in node])
        if isinstance(node, dict):
            keys = sorted(node)
            return (dict, keys, [rec(node[k]) for k in keys])
        leaves.append(node)
        return None

    return leaves, rec(tree)


def unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(td):
        if td is None:
            return next(it)
        if td[0] is dict:
            return {k: rec(c) for k, c in zip(td[1], td[2])}
        seq = [rec(c) for c in td[1]]
        return seq if td[0] is list else td[0](seq)

    return rec(treedef)
