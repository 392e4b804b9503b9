"""H3WB 134-joint skeleton tables (integer, bit-exact) without the dataset files.

The reference builds these tables inside ``Human3WBDataset.__init__`` from the
``metadata`` dict stored in ``data/train_h3wb.npz`` (absent here), see
``common/h3wb_dataset.py:26-61`` (left/right lists, root_indices,
parts_connection_indices) and ``common/h3wb_dataset.py:198-213``
(parts_joint_indices).  The H3WB / COCO-WholeBody keypoint order is fixed
(17 body, 3+3 feet, 68 face, 21+21 hands = 133, plus the synthetic root the
reference prepends, ``h3wb_dataset.py:163-193``), so the tables are restated
here as constants.  ``H3WBSkeleton`` exposes exactly the attributes the hot
path reads from ``dataset`` (``common/diffusionpose.py:72-75``,
``common/utils.py:113-126``).

UNVERIFIED AGAINST THE DATASET FILE: the group boundaries follow from the fixed H3WB keypoint order, but the
left/right symmetry lists (``metadata['left_side']`` / ``['right_side']``, including the reference's duplicate
filtering at ``h3wb_dataset.py:29-38``) live in the absent npz; the lists below are the anatomical pairs of the
COCO-WholeBody layout and have not been compared with the file.  They only enter through ``joints_left`` /
``joints_right``, which ``D3DP`` takes as constructor arguments: with the real dataset, pass the dataset's own lists
(``H3WBSkeleton.from_metadata(np.load('train_h3wb.npz', allow_pickle=True)['metadata'].item())``) -- flip-TTA parity
with a real checkpoint depends on them.
"""
from __future__ import annotations

import copy

NUM_KPS = 134  # 133 H3WB keypoints + prepended root (config.yaml:20)

# 0-based ids inside the 133-keypoint H3WB layout (== metadata[...] of the npz)
_META_133 = {
    "body": list(range(0, 17)),
    "left_foot": list(range(17, 20)),
    "right_foot": list(range(20, 23)),
    "face": list(range(23, 91)),
    "left_hand": list(range(91, 112)),
    "right_hand": list(range(112, 133)),
}


def _face68_pairs():
    """Left/right pairs of the 68-landmark face layout (iBUG order)."""
    pairs = [(i, 16 - i) for i in range(8)]                 # jaw line, 8 is the chin
    pairs += [(17 + i, 26 - i) for i in range(5)]           # eyebrows
    pairs += [(31, 35), (32, 34)]                           # nostrils, 27-30 and 33 are central
    pairs += [(36, 45), (37, 44), (38, 43), (39, 42), (40, 47), (41, 46)]  # eyes
    pairs += [(48, 54), (49, 53), (50, 52), (59, 55), (58, 56)]            # outer lip
    pairs += [(60, 64), (61, 63), (67, 65)]                                # inner lip
    return pairs


def _symmetry_133():
    left, right = [], []
    # COCO body: odd ids are the subject's left side, even ids (>0) the right side
    for i in range(1, 17, 2):
        left.append(i)
        right.append(i + 1)
    left += _META_133["left_foot"]
    right += _META_133["right_foot"]
    for a, b in _face68_pairs():
        # image-left landmarks (low ids) belong to the subject's right side
        right.append(23 + a)
        left.append(23 + b)
    left += _META_133["left_hand"]
    right += _META_133["right_hand"]
    return left, right


class H3WBSkeleton:
    """Stand-in for ``Human3WBDataset`` carrying only the hot-path tables.

    Attributes mirror the reference object: ``metadata``, ``root_indices``,
    ``parts_connection_indices``, ``parts_joint_indices``,
    ``keypoints_metadata`` (``h3wb_dataset.py:49-67,198-213``).
    """

    def __init__(self, add_root: bool = True):
        offset = 1 if add_root else 0
        left, right = _symmetry_133()
        self.metadata = copy.deepcopy(_META_133)
        self.metadata["left_side"] = list(left)
        self.metadata["right_side"] = list(right)
        self.joints_left = [j + offset for j in left]
        self.joints_right = [j + offset for j in right]
        self.num_kps = 133 + offset
        self.root_indices = {"body": 0, "face": 54, "left_hand": 92, "right_hand": 113}
        self.parts_connection_indices = {"face": 1, "left_hand": 10, "right_hand": 11}
        self.keypoints_metadata = {
            "layout_name": "h3wb",
            "num_joints": self.num_kps,
            "keypoints_symmetry": [self.joints_left, self.joints_right],
        }
        pji = {p: [j + 1 for j in self.metadata[p]]
               for p in ("body", "face", "left_hand", "right_hand", "left_foot", "right_foot")}
        pji["body"] = [0] + pji["body"] + pji["left_foot"] + pji["right_foot"]
        del pji["left_foot"], pji["right_foot"]
        self.parts_joint_indices = pji

    @classmethod
    def from_metadata(cls, metadata: dict, add_root: bool = True):
        """Skeleton whose left/right side lists come from the ``metadata`` dict of ``train_h3wb.npz``
        (``h3wb_dataset.py:26-38``, 0-based ids of the 133-keypoint layout) instead of the built-in constants; whether
        the two agree is recorded in ``symmetry_matches_builtin`` (if not, flip-TTA pairs differ from the ones the
        synthetic tests exercised -- the kernels take the permutation as data, so the path itself is unaffected)."""
        sk = cls(add_root=add_root)
        offset = 1 if add_root else 0
        left, right = list(metadata["left_side"]), list(metadata["right_side"])
        # h3wb_dataset.py:29-38: a joint listed on both sides is dropped from both lists
        dups = [kp for kp in left if kp in right]
        left = [int(e) for e in left if e not in dups]
        right = [int(e) for e in right if e not in dups]
        sk.symmetry_matches_builtin = (sorted(zip(left, right)) == sorted(zip(sk.metadata["left_side"], sk.metadata["right_side"])))
        sk.metadata["left_side"], sk.metadata["right_side"] = left, right
        sk.joints_left = [int(j) + offset for j in left]
        sk.joints_right = [int(j) + offset for j in right]
        sk.keypoints_metadata["keypoints_symmetry"] = [sk.joints_left, sk.joints_right]
        return sk

    def kps_left(self):
        return list(self.joints_left)

    def kps_right(self):
        return list(self.joints_right)


def merged_part_indices(parts_joint_indices, merge_hands=True):
    """``D3DP.__init__`` merge of the two hand groups (``diffusionpose.py:76-83``)."""
    pji = {k: list(v) for k, v in parts_joint_indices.items()}
    if merge_hands:
        pji["hands"] = pji["left_hand"] + pji["right_hand"]
        del pji["left_hand"], pji["right_hand"]
    return pji


def flip_permutation(joints_left, joints_right, num_kps=NUM_KPS):
    """perm[j] = source joint of j under the left/right swap used by flip-TTA
    (``x[..., L+R, :] = x[..., R+L, :]``, ``diffusionpose.py:197-198``)."""
    perm = list(range(num_kps))
    for l, r in zip(joints_left, joints_right):
        perm[l] = r
        perm[r] = l
    return perm
