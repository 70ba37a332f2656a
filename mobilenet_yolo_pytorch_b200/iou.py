"""Drop-in for the reference's ``utils/iou.py``: broadcast (n1, n2) box overlap on
xyxy boxes, one kernel each instead of ~10 elementwise launches with
(n1, n2, 2) temporaries."""
from __future__ import annotations

from . import ops


def find_intersection(set_1, set_2):
    """utils/iou.py:4-13"""
    return ops.pairwise(set_1, set_2, 0)


def find_union(set_1, set_2):
    """utils/iou.py:14-31"""
    return ops.pairwise(set_1, set_2, 1)


def find_jaccard_overlap(set_1, set_2):
    """utils/iou.py:32-49"""
    return ops.pairwise(set_1, set_2, 2)
