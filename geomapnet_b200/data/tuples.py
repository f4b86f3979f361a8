"""Index arithmetic of the reference's multi-frame datasets (host side; the frames themselves are gathered on the
GPU by ImagePipeline).

Mirror of /root/reference/dataset_loaders/composite.py:
 * ``MF.get_indices`` (:60-75): tuple `index` -> the `steps` frame indices it is made of: offsets 0, skip, 2 skip, ...
   centred on their middle element (py2 integer division), shifted by (steps // 2) * skip when ``no_duplicates``,
   clipped to [0, L-1]; ``variable_skip`` draws every gap from 1..skip (np.random, as the reference);
 * ``MF.__len__`` (:105-109);
 * ``MFOnline.__getitem__`` (:117-126): T frames of the training sequence (absolute poses) followed by T frames of the
   validation sequence (no_duplicates, relative poses = VOs supplied by the dataset), both addressed modulo their length.
Targets are gathered from pose tables with the same indices; ``calc_vos_simple`` (the reference's train-time vo_func,
common/pose_utils.py:234-246) is a plain difference.  The numpy ``calc_vos_safe`` the reference uses for MFOnline's
validation half operates on the dataset's DSO/SLAM poses and stays with the dataset code (out of scope).
"""
import numpy as np

__all__ = ["mf_offsets", "mf_indices", "mf_len", "mfonline_len", "mfonline_indices", "batch_frame_indices"]


def mf_offsets(steps, skip, no_duplicates=False, variable_skip=False, rng=None):
    if variable_skip:
        rng = np.random if rng is None else rng
        skips = rng.randint(1, high=skip + 1, size=steps - 1)
    else:
        skips = skip * np.ones(steps - 1)
    offsets = np.insert(skips, 0, 0).cumsum()
    offsets = offsets - offsets[len(offsets) // 2]
    if no_duplicates:
        offsets = offsets + (steps // 2) * skip
    return offsets.astype(np.int64)


def mf_indices(index, n_frames, steps, skip, no_duplicates=False, variable_skip=False, rng=None):
    """composite.py:60-75 -> int64 [steps]"""
    idx = index + mf_offsets(steps, skip, no_duplicates, variable_skip, rng)
    return np.minimum(np.maximum(idx, 0), n_frames - 1)


def mf_len(n_frames, steps, skip, no_duplicates=False):
    """composite.py:105-109"""
    return n_frames - (steps - 1) * skip if no_duplicates else n_frames


def mfonline_len(n_val_frames, steps, skip):
    """composite.py:128-129: the length of the validation MF (no_duplicates=True)"""
    return mf_len(n_val_frames, steps, skip, no_duplicates=True)


def mfonline_indices(idx, n_train_frames, n_val_frames, steps, skip, variable_skip=False, rng=None):
    """composite.py:117-126 -> (train frame indices [steps], validation frame indices [steps])"""
    train_idx = idx % mf_len(n_train_frames, steps, skip, False)
    val_idx = idx % mf_len(n_val_frames, steps, skip, True)
    return (mf_indices(train_idx, n_train_frames, steps, skip, False, variable_skip, rng),
            mf_indices(val_idx, n_val_frames, steps, skip, True, variable_skip, rng))


def batch_frame_indices(tuple_indices, n_frames, steps, skip, no_duplicates=False):
    """A minibatch of MF tuples -> int32 [N * steps] frame indices in the [N, T] order MapNet.forward folds into the
    batch (models/posenet.py:93-97): what ImagePipeline(frames, index=...) takes."""
    out = np.stack([mf_indices(int(i), n_frames, steps, skip, no_duplicates) for i in tuple_indices], 0)
    return out.reshape(-1).astype(np.int32)
