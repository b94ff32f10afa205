"""Minimal stand-in for the reference's batch container (recbole/data/interaction.py:19-347).

The product path accepts ANY mapping ``interaction[field] -> Tensor`` (the reference's own
``Interaction`` included); this class only exists so that the package can be driven without the
reference installed (tests, bench, the GPU box).  It mirrors the members the hot path touches:
``__getitem__``, ``__len__``/``length``, ``to(device)`` (which, like the reference's
interaction.py:138-164, does NOT carry pos_len_list/user_len_list over) and
``set_additional_info``.
"""
import torch


class Interaction:
    def __init__(self, interaction, pos_len_list=None, user_len_list=None):
        self.interaction = dict(interaction)
        self.pos_len_list = pos_len_list
        self.user_len_list = user_len_list
        self.length = -1
        for v in self.interaction.values():
            self.length = max(self.length, v.shape[0])

    def set_additional_info(self, pos_len_list=None, user_len_list=None):
        self.pos_len_list = pos_len_list
        self.user_len_list = user_len_list

    def __getitem__(self, index):
        if isinstance(index, str):
            return self.interaction[index]
        return Interaction({k: v[index] for k, v in self.interaction.items()})

    def __contains__(self, item):
        return item in self.interaction

    def __len__(self):
        return self.length

    def __iter__(self):
        return iter(self.interaction)

    def to(self, device, selected_field=None):
        sel = set(self.interaction) if selected_field is None else set(
            [selected_field] if isinstance(selected_field, str) else selected_field)
        return Interaction({k: (v.to(device, non_blocking=True) if k in sel else v)
                            for k, v in self.interaction.items()})

    def cpu(self):
        return Interaction({k: v.cpu() for k, v in self.interaction.items()})
