"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the LLM-prompt assembly of the reference.

Follows ``models/xinstructblip.py:342-386`` (generate) and ``:544-594`` (forward): for every frame position the optional
enumeration tokens, then for modality in ['video', 'audio'] the cue embeddings (repeated over the batch) and that frame's
32 projected query tokens (``inputs_llm[m].view(bs, num, 32, -1)[:, pos]``), then the frame's timestamp tokens; after the
loop the duration tokens and the prompt tokens; ``torch.cat(..., dim=1)`` of embeddings and of attention masks;
``empty_targets`` (-100) over the multimodal prefix (:579-588).  The LLM embedding lookups themselves are inputs (the LLM
is frozen and outside the hot path).  Pinned by construction: it is the same sequence of ``torch.cat`` inputs as the
reference lines cited.  Only tests / smoke may import this module.
"""
import torch


def assemble(inputs_llm, atts_llm, cue_embeds, cue_atts, duration_embeds, duration_atts, prompt_embeds, prompt_atts,
             timestamp_embeds=None, timestamp_atts=None, enumeration_embeds=None, num_query_token=32):
    """inputs_llm[m]: [bs, F*32, D]; returns (inputs_embeds [bs, L, D], attention_mask [bs, L])."""
    mods = [m for m in ("video", "audio") if m in inputs_llm]     # :359 / :560 hard-coded order
    bs = prompt_embeds.shape[0]
    num = {m: inputs_llm[m].shape[1] // num_query_token for m in mods}
    att_list, inp_list = [], []
    for pos in range(num[mods[0]]):
        if enumeration_embeds is not None:                         # :349-357
            e = enumeration_embeds[pos]
            inp_list.append(e.unsqueeze(0).repeat(bs, 1, 1))
            att_list.append(torch.ones(bs, e.shape[0], dtype=torch.long))
        for m in mods:                                             # :360-362
            att_list.extend([cue_atts[m].view(1, -1).repeat(bs, 1),
                             atts_llm[m].view(bs, num[m], num_query_token)[:, pos, :]])
            inp_list.extend([cue_embeds[m].unsqueeze(0).repeat(bs, 1, 1),
                             inputs_llm[m].view(bs, num[m], num_query_token, -1)[:, pos, :, :]])
        if timestamp_embeds is not None:                           # :364-366
            inp_list.append(timestamp_embeds[:, pos, :, :])
            att_list.append(timestamp_atts[:, pos, :])
    att_list.append(duration_atts)                                 # :369-378
    inp_list.append(duration_embeds)
    prefix_len = sum(a.shape[1] for a in att_list)
    att_list.append(prompt_atts)                                   # :381-383
    inp_list.append(prompt_embeds)
    return torch.cat(inp_list, dim=1), torch.cat(att_list, dim=1), prefix_len


def targets(text_targets, prefix_len):
    """:579-588"""
    empty = torch.ones(text_targets.shape[0], prefix_len, dtype=torch.long).fill_(-100)
    return torch.cat([empty, text_targets], dim=1)
