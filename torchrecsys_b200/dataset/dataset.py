"""Host-side data preparation with the reference's interface (dataset/dataset.py).

``ProcessData`` turns the interactions frame into dictionaries of LongTensors and
``FastDataLoader`` iterates them batch by batch; both are pure host objects that work without a
GPU (the reference tests use them that way).  Differences from the reference, all supersets:

* metadata columns may hold scalar ints, python lists or str-encoded lists (the reference only
  survives the last); each (row, feature) is reduced to ONE id -- bag element 0,
  the only element the reference scorers read (D7) -- so ``pos_metadata_id`` is ``[N, F]``.
* everything is vectorised (no ``iterrows`` / per-sample python loops), and dynamic negatives use
  a vectorised rejection loop with the same contract: uniform over items, never equal to the
  positive (dataset.py:435-447).
* ``item_meta``: dense ``[n_items, F]`` table of an item's metadata ids (the array form of
  ``item_to_metadata_map``), used by the on-device negative sampler.

``fit`` itself does not iterate this loader: it keeps the split on the device and samples /
shuffles there (see engine.py)."""
from __future__ import annotations

import ast
import json
from typing import Dict, List, Optional

import numpy as np
import pandas as pd
import torch
from sklearn.model_selection import train_test_split


def _first_id(x) -> int:
    """Bag element 0 of a metadata cell; 0 (the reference's pad value) for an empty bag."""
    if isinstance(x, str):
        try:
            x = ast.literal_eval(x)
        except (ValueError, SyntaxError):
            return 0
    if isinstance(x, (list, tuple, np.ndarray)):
        return int(x[0]) if len(x) else 0
    if x is None or (isinstance(x, float) and np.isnan(x)):
        return 0
    return int(x)


def _as_bag(x) -> list:
    if isinstance(x, str):
        try:
            x = ast.literal_eval(x)
        except (ValueError, SyntaxError):
            return []
    if isinstance(x, (list, tuple, np.ndarray)):
        return [int(v) for v in x]
    if x is None or (isinstance(x, float) and np.isnan(x)):
        return []
    return [int(x)]


class Data:
    """Counts users / items / metadata categories and draws the static negatives
    (reference dataset.py:15-64)."""

    def __init__(self, dataset: pd.DataFrame, user_id_col: str, item_id_col: str,
                 metadata_id_col: Optional[List[str]] = None, split_ratio: float = 0.8,
                 dynamic_neg_sampling: bool = False):
        self.dataset = dataset
        self.user_id, self.item_id = user_id_col, item_id_col
        self.num_items = int(dataset[item_id_col].nunique())
        self.num_users = int(dataset[user_id_col].nunique())
        self.dynamic_neg_sampling = dynamic_neg_sampling
        self.split_ratio = split_ratio
        self._neg_items = None
        if not dynamic_neg_sampling:
            # same draw as the reference: numpy global RNG, may coincide with the positive
            self._neg_items = np.random.randint(low=0, high=self.num_items, size=len(dataset))
        if metadata_id_col:
            self.metadata_id = list(metadata_id_col)
            self.negative_metadata_id = ["neg_" + c for c in self.metadata_id]
            self._meta_first = {c: np.fromiter((_first_id(v) for v in dataset[c]), dtype=np.int64,
                                               count=len(dataset)) for c in self.metadata_id}
            self.metadata_size = {}
            for c in self.metadata_id:
                cells = dataset[c]
                try:
                    self.metadata_size[c] = int(cells.nunique())
                except TypeError:  # unhashable list cells
                    self.metadata_size[c] = int(cells.map(repr).nunique())
                # every id that can be looked up must have a row
                self.metadata_size[c] = max(self.metadata_size[c], int(self._meta_first[c].max()) + 1)


class ProcessData(Data):
    def __init__(self, dataset: pd.DataFrame, user_id_col: str, item_id_col: str,
                 metadata_id_col: Optional[List[str]] = None, split_ratio: float = 0.9,
                 dynamic_neg_sampling: bool = False):
        super().__init__(dataset, user_id_col, item_id_col, metadata_id_col, split_ratio,
                         dynamic_neg_sampling)
        self.user_id_col, self.item_id_col = user_id_col, item_id_col
        if metadata_id_col:
            self.metadata_id_col = list(metadata_id_col)

    # -------------------------------------------------------------------------------------
    def prepare_data(self) -> None:
        """Builds ``train_data`` / ``test_data`` (dicts of CPU LongTensors), ``config``,
        ``item_to_metadata_map``, ``meta_data_df`` (reference dataset.py:140-249)."""
        cols = {"user_id": self.dataset[self.user_id_col].to_numpy(dtype=np.int64),
                "pos_item_id": self.dataset[self.item_id_col].to_numpy(dtype=np.int64)}
        if not self.dynamic_neg_sampling:
            cols["neg_item_id"] = np.asarray(self._neg_items, dtype=np.int64)
        meta_cols = getattr(self, "metadata_id_col", None)
        self.item_meta = None
        self.item_to_metadata_map = None
        self.meta_data_df = None
        if meta_cols:
            pos_meta = np.stack([self._meta_first[c] for c in meta_cols], axis=1)
            # dense item -> metadata table: the first row seen for each item wins
            item_meta = np.zeros((self.num_items, len(meta_cols)), dtype=np.int64)
            items = cols["pos_item_id"]
            uniq, first = np.unique(items, return_index=True)
            ok = (uniq >= 0) & (uniq < self.num_items)
            item_meta[uniq[ok]] = pos_meta[first[ok]]
            self.item_meta = item_meta
            cols["pos_metadata_id"] = pos_meta
            if not self.dynamic_neg_sampling:
                neg = np.clip(cols["neg_item_id"], 0, self.num_items - 1)
                cols["neg_metadata_id"] = item_meta[neg]
            bags = {c: self.dataset[c].iloc[first].map(_as_bag).tolist() for c in meta_cols}
            self.item_to_metadata_map = {int(it): {c: bags[c][k] for c in meta_cols}
                                         for k, it in enumerate(uniq)}
            self.meta_data_df = pd.DataFrame({"pos_item_id": uniq, **{c: bags[c] for c in meta_cols}})

        self.config = {"num_users": self.num_users, "num_items": self.num_items,
                       "num_metadata": getattr(self, "metadata_size", {})}

        n = len(self.dataset)
        if self.split_ratio < 1 and n > 1:
            # identical call to the reference's (dataset.py:240) on a row index -> identical split
            tr, te = train_test_split(np.arange(n), test_size=1 - self.split_ratio, random_state=42)
        else:
            tr, te = np.arange(n), np.arange(0)
        self.train_data = self._to_tensor_dict(cols, tr)
        self.test_data = self._to_tensor_dict(cols, te)

    @staticmethod
    def _to_tensor_dict(cols: Dict[str, np.ndarray], rows: np.ndarray) -> Dict[str, torch.Tensor]:
        return {k: torch.from_numpy(np.ascontiguousarray(v[rows])).long() for k, v in cols.items()}

    def write_data(self, path: str) -> None:
        """config.json and meta.csv, as the reference (dataset.py:307-316)."""
        with open(f"{path}/config.json", "w") as fh:
            json.dump(self.config, fh)
        if self.meta_data_df is not None:
            self.meta_data_df.to_csv(f"{path}/meta.csv", index=False)


class FastDataLoader:
    """In-process batch iterator over a dict of tensors (reference dataset.py:319-458)."""

    _POS_KEYS = ("user_id", "pos_item_id", "pos_metadata_id")
    _NEG_KEYS = ("neg_item_id", "neg_metadata_id")

    def __init__(self, data: dict, batch_size: int = 32, shuffle: bool = False,
                 dynamic_neg_sampling: bool = False, n_items: Optional[int] = None,
                 item_to_metadata_map: Optional[dict] = None,
                 metadata_id_cols: Optional[List[str]] = None):
        if dynamic_neg_sampling and n_items is None:
            raise ValueError("n_items must be provided for dynamic negative sampling.")
        if dynamic_neg_sampling and metadata_id_cols and item_to_metadata_map is None:
            raise ValueError("item_to_metadata_map must be provided for dynamic negative sampling with metadata.")
        self.data, self.batch_size, self.shuffle = data, batch_size, shuffle
        self.dynamic_neg_sampling, self.n_items = dynamic_neg_sampling, n_items
        self.item_to_metadata_map, self.metadata_id_cols = item_to_metadata_map, metadata_id_cols
        user = data.get("user_id")
        self.dataset_len = int(user.shape[0]) if torch.is_tensor(user) else 0
        self.num_batches = -(-self.dataset_len // batch_size) if self.dataset_len else 0
        self._item_meta = None
        if dynamic_neg_sampling and metadata_id_cols and item_to_metadata_map:
            table = np.zeros((int(n_items), len(metadata_id_cols)), dtype=np.int64)
            for it, feats in item_to_metadata_map.items():
                if 0 <= int(it) < n_items:
                    for f, c in enumerate(metadata_id_cols):
                        bag = feats.get(c, [])
                        table[int(it), f] = bag[0] if len(bag) else 0
            self._item_meta = torch.from_numpy(table)
        if shuffle and self.dataset_len:
            self.shuffle_indices()

    def shuffle_indices(self) -> None:
        if self.dataset_len:
            self.indices = torch.randperm(self.dataset_len)

    def __len__(self) -> int:
        return self.num_batches

    def __iter__(self):
        self.i = 0
        if self.shuffle and self.dataset_len:
            self.shuffle_indices()
        return self

    def __next__(self):
        if self.i >= self.dataset_len:
            raise StopIteration
        end = min(self.i + self.batch_size, self.dataset_len)
        sel = self.indices[self.i:end] if (self.shuffle and self.dataset_len) else slice(self.i, end)
        keys = self._POS_KEYS + (() if self.dynamic_neg_sampling else self._NEG_KEYS)
        batch = {k: self.data[k][sel] for k in keys if k in self.data}
        if self.dynamic_neg_sampling:
            pos = batch["pos_item_id"].numpy()
            neg = np.random.randint(0, self.n_items, size=pos.shape[0])
            clash = neg == pos
            while clash.any():  # redraw only the collisions
                neg[clash] = np.random.randint(0, self.n_items, size=int(clash.sum()))
                clash = neg == pos
            batch["neg_item_id"] = torch.from_numpy(neg).long()
            if self._item_meta is not None and "pos_metadata_id" in batch:
                batch["neg_metadata_id"] = self._item_meta[batch["neg_item_id"]]
        self.i += self.batch_size
        return batch
