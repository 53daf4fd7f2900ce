"""Data-parallel sharding of independent prompts over the GPUs of one box.

The reference has no multi-device code; its only batching is left-padding prompts in one process
(/root/reference/phi.py:236-245). Prompts are independent, so the B200 build shards them: one
process per GPU, weights replicated, NO collective on the data path. torch.distributed is used
only by the harness (gathering the finished rows to rank 0, barriers around timed regions).

H7 (static LongRoPE switch, phi.py:492,583): the reference picks short/long factors from the
padded length of the *whole* batch, so every shard must use the switch implied by the global
maximum length — `global_rope_switch` computes it and the model takes it as `force_long_rope`.
"""
import torch
import torch.distributed as dist


def shard_indices(lengths, world):
    """Length-sorted round-robin: rank r gets sorted_idx[r::world]. Keeps per-shard padding small
    and the per-rank token counts balanced. Returns a list (per rank) of original indices."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    return [order[r::world] for r in range(world)]


def global_rope_switch(lengths, max_tokens, original_max=4096):
    return (max(lengths) + max_tokens) > original_max if lengths else False


def dp_map(fn, items, lengths=None, group=None):
    """Run fn(shard_items, shard_indices) on this rank's shard and return, on rank 0, the results
    re-ordered to input order (other ranks get None). Works without an initialised process group
    (single process). fn must return one result per item."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lengths = lengths if lengths is not None else [1] * len(items)
    shards = shard_indices(lengths, world)
    mine = shards[rank]
    res = fn([items[i] for i in mine], mine) if mine else []
    if len(res) != len(mine):
        raise ValueError('dp_map: fn must return one result per item')
    if world == 1:
        out = [None] * len(items)
        for i, r in zip(mine, res):
            out[i] = r
        return out
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(list(zip(mine, res)), gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * len(items)
    for part in gathered:
        for i, r in part:
            out[i] = r
    return out


def dp_generate(model, processor, prompts, images=None, max_tokens=512, apply_chat_template=False, **kw):
    """generate() over a prompt list sharded across ranks; rank 0 returns the texts in input order (other ranks None).
    `images`: None (text-only: the reference's left-padded batch, phi:236-245) or one entry per prompt (None | image | list of
    images): each rank then runs its shard through api.generate_batch (per-prompt batch-1 semantics, H11)."""
    from .api import _generate, _prompt_lengths, _apply_chat_template, generate_batch
    prompts = list(prompts)
    if images is not None:
        if len(images) != len(prompts):
            raise ValueError('images must hold one entry per prompt')
        norm = [None if im is None else (list(im) if isinstance(im, (list, tuple)) else [im]) for im in images]
        texts = [_apply_chat_template(p, im, False, apply_chat_template)[0] for p, im in zip(prompts, norm)]
        lens = _prompt_lengths(processor, texts, norm)
    else:
        lens = [len(processor.tokenizer(p).input_ids) for p in prompts]
    prev = model.force_long_rope
    model.force_long_rope = global_rope_switch(lens, max_tokens, model.cfg.original_max_position_embeddings)

    def run(shard, idx):
        if images is not None:
            return generate_batch(shard, [images[i] for i in idx], preload=(model, processor), max_tokens=max_tokens,
                                  verbose=False, apply_chat_template=apply_chat_template, **kw)
        out = _generate(model, processor, shard if len(shard) > 1 else shard[0], max_tokens=max_tokens, verbose=False,
                        stream=False, mute=True, **kw)
        return out if isinstance(out, list) else [out]
    try:
        return dp_map(run, prompts, lens)
    finally:
        model.force_long_rope = prev
