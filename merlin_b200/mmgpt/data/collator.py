"""DataCollatorForSupervisedDataset — mirror of mmgpt/data/collator.py:8-34: right-pads `input_ids` with the
tokenizer's pad id and `labels` with IGNORE_INDEX, truncates both to tokenizer.model_max_length, derives
`attention_mask = input_ids != pad`, and stacks each sample's image list into one [n_i, 3, H, W] tensor.  The
resulting dict is exactly what MMGPTLlamaForCausalLM.forward / MMGPTTrainer.training_step consume."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from ..utils.constants import IGNORE_INDEX


@dataclass
class DataCollatorForSupervisedDataset(object):
    tokenizer: object

    def __call__(self, instances):
        input_ids, labels = tuple([instance[key] for instance in instances] for key in ("input_ids", "labels"))
        images = [torch.stack(instance["image"]) for instance in instances]
        input_ids = torch.nn.utils.rnn.pad_sequence(input_ids, batch_first=True,
                                                    padding_value=self.tokenizer.pad_token_id)
        labels = torch.nn.utils.rnn.pad_sequence(labels, batch_first=True, padding_value=IGNORE_INDEX)
        input_ids = input_ids[:, :self.tokenizer.model_max_length]
        labels = labels[:, :self.tokenizer.model_max_length]
        return dict(input_ids=input_ids, labels=labels, attention_mask=input_ids.ne(self.tokenizer.pad_token_id),
                    images=images)
