"""Concept bookkeeping of model/text_embeddings.py:188-215 (the part the hot path uses).

The reference derives class -> concept-row indices from its concept name lists; only the group sizes matter here
(voc12_wbg_concept4_single.npy has 98 rows for 21 classes, cityscapes_concept3_single.npy 54 rows for 19 classes)."""
import os

import torch

_CONCEPT_COUNTS = {
    "voc12_wbg_concept4_single.npy": [45, 3, 3, 1, 4, 4, 2, 4, 2, 3, 1, 2, 2, 4, 4, 4, 3, 1, 1, 2, 3],
    "cityscapes_concept3_single.npy": [3, 1, 7, 1, 2, 3, 1, 3, 3, 4, 1, 7, 3, 4, 4, 1, 2, 3, 1],
}


def get_class_to_concept_idxs(save_path):
    counts = _CONCEPT_COUNTS.get(os.path.basename(save_path))
    if counts is None:
        raise ValueError(save_path)
    out, k = {}, 0
    for i, c in enumerate(counts):
        out[i] = list(range(k, k + c))
        k += c
    return out


def concept_offsets(save_path, num_rows, num_classes, device):
    """int32 [num_classes + 1] prefix offsets for svl_group_max (identity when the table has one row per class)."""
    if num_rows == num_classes:
        return torch.arange(num_classes + 1, device=device, dtype=torch.int32)
    groups = get_class_to_concept_idxs(save_path)
    assert len(groups) == num_classes and sum(len(g) for g in groups.values()) == num_rows
    offs = [0]
    for i in range(num_classes):
        offs.append(offs[-1] + len(groups[i]))
    return torch.tensor(offs, device=device, dtype=torch.int32)


def aggregate_concept_predictions(pred, class_to_concept_idxs):
    """max over the concept channels of each class (model/text_embeddings.py:188-193); pred [B, K, H, W] -> [B, N, H, W]."""
    return torch.stack([pred[:, idx].max(dim=1).values for _, idx in sorted(class_to_concept_idxs.items())], dim=1)
