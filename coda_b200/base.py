"""Selector protocol shared by every acquisition strategy.

The drivers (reference ``main.py:91-94``, ``demo/app.py``) only ever make three calls on a
selector; this abstract base pins their signatures (reference ``coda/base.py:1-16``):

    idx, prob = selector.get_next_item_to_label()
    selector.add_label(idx, true_class, prob)
    best = selector.get_best_model_prediction()
"""
from __future__ import annotations

import abc


class ModelSelector(abc.ABC):
    """Active model-selection strategy over an (H models, N items, C classes) prediction slab."""

    @abc.abstractmethod
    def get_next_item_to_label(self):
        """Pick the item to send to the annotator -> (global item index, acquisition value)."""

    @abc.abstractmethod
    def add_label(self, chosen_idx, true_class, selection_prob):
        """Fold the revealed class of ``chosen_idx`` into the selector's posterior."""

    @abc.abstractmethod
    def get_best_model_prediction(self):
        """Index of the model currently believed to be best."""
