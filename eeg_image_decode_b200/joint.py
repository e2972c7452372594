"""Joint-subject variant of ATM-S behind the reference's surface.

Drop-in for the classes/functions of Retrieval/ATMS_retrieval_joint_train.py: ``ATMS(sequence_length=250,
num_subjects=10, joint_train=False)`` (:172-191) whose DataEmbedding holds one value embedding (Linear 250->250) per
subject when ``joint_train=True`` (models/subject_layers/Embed.py:127-130, 144), and the script's ``train_model`` /
``evaluate_model`` / ``main_train_loop`` (:201-254, :257-359, :361-...), which have the same bodies as the single-subject
script.  The reference evaluates ``self.value_embedding[str(subject_id.item())](x[i])`` per trial in a Python loop with a
device->host sync each; here the batch is ordered by subject and every subject present runs as ONE tcgen05 GEMM over
its trials (``eegb200_atms_forward`` with ``joint_value_w`` / ``group_offsets``, include/eegdecode_b200.h), the weight
gradient likewise.  Reference behaviour kept on purpose:
  * an id without a value embedding (>= 10, e.g. ``sub-10``) raises ``KeyError`` like the ModuleDict lookup does;
  * only the value embeddings of subjects present in the batch receive a gradient; AdamW leaves the others untouched
    (no weight decay either), as torch.optim does for ``grad is None``;
  * state_dict keys ``encoder.enc_embedding.value_embedding.<s>.{weight,bias}`` and ``num_subjects`` subject_wise_linear
    layers, so checkpoints interchange with the reference (``strict=True``).
"""
from __future__ import annotations

from .atms import ATMS as _ATMSBase
from .train import evaluate_model, extract_id_from_string, main_train_loop, train_model  # noqa: F401  (same bodies)


class ATMS(_ATMSBase):
    """``ATMS(sequence_length=250, num_subjects=10, joint_train=False)`` (ATMS_retrieval_joint_train.py:173)"""

    def __init__(self, sequence_length=250, num_subjects=10, joint_train=False):
        super().__init__(num_channels=63, sequence_length=sequence_length, num_subjects=num_subjects,
                         _joint_train=joint_train)
