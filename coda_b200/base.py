"""The selector interface the drivers call (reference coda/base.py:1-16)."""


class ModelSelector:
    def __init__(self):
        pass

    def get_next_item_to_label(self):
        """Return (index, selection probability)."""
        raise NotImplementedError

    def add_label(self, chosen_idx, true_class, selection_prob):
        raise NotImplementedError

    def get_best_model_prediction(self):
        raise NotImplementedError
