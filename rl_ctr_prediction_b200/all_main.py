"""RL + CTR pipeline step -- drop-in for the loop functions of the reference's ``src/all_main/main.py``
(``get_model`` :20-30 region, ``generate_preds`` :183-271, ``train`` :274-319, ``test`` :321-347).

Per batch: frozen state encoder (rlctr_featemb_fwd) -> DDQN picks how many CTR models
to ensemble -> DDPG actor emits their weights -> generate_preds scores the M frozen CTR models (M fused
gather+interaction kernels writing one [B, M] buffer, then rlctr_generate_preds) and computes the +-1
reward -> transitions go to the device ring buffers -> one DDQN learn step and one DDPG critic/actor
learn step on replayed batches (tcgen05 GEMMs + fused dense Adam) -> Polyak updates.
"""
from __future__ import annotations

import torch

from . import DDPG_for_PG_model, DDQN_model
from .ensemble import generate_preds  # noqa: F401  (same name and signature as main.py:183)


def get_model(action_nums, feature_nums, field_nums, latent_dims, batch_size, memory_size, device, campaign_id):
    """main.py: builds the two agents of the pipeline."""
    ddqn = DDQN_model.DoubleDQN(feature_nums, field_nums, latent_dims, action_nums=action_nums, campaign_id=campaign_id,
                                batch_size=batch_size, memory_size=memory_size, device=device)
    ddpg = DDPG_for_PG_model.DDPG(feature_nums, field_nums, latent_dims, action_nums=action_nums, campaign_id=campaign_id,
                                  batch_size=batch_size, memory_size=memory_size, device=device)
    return ddqn, ddpg


def train_step(ddqn_model, ddpg_for_pg_model, model_dict, features, labels, embedding_layer, exploration_rate, device):
    """Body of train() at main.py:281-317 for one batch.  Returns (y_preds, rewards, td_error, a_loss)."""
    embedding_vectors = embedding_layer.forward(features)
    actions = ddqn_model.choose_action(embedding_vectors, exploration_rate)
    prob_weights = ddpg_for_pg_model.choose_action(embedding_vectors, actions.float(), exploration_rate)
    y_preds, prob_weights_new, rewards = generate_preds(model_dict, features, actions, prob_weights, labels, device,
                                                        mode="train")
    ddqn_model.store_transition(torch.cat([features, actions, rewards.long()], dim=1))
    ddpg_for_pg_model.store_transition(features, torch.cat([prob_weights_new, rewards], dim=1), actions.float())
    b_s, b_a, b_r, b_s_ = ddqn_model.sample_batch()
    gather = getattr(ddqn_model, "batch_gather", None)       # data-parallel run (make_gathered_replay): the ranks' samples joined
    if gather is not None:
        b_s, b_a, b_r = gather(b_s, b_a, b_r)
        b_s_ = b_s
    ddqn_model.learn(embedding_layer.forward(b_s), b_a, b_r, embedding_layer.forward(b_s_))
    b_s, b_a, b_r, b_s_, b_pg_a = ddpg_for_pg_model.sample_batch()
    if gather is not None:
        b_s, b_a, b_r, b_pg_a = gather(b_s, b_a, b_r, b_pg_a)
        b_s_ = b_s
    es, es_ = embedding_layer.forward(b_s), embedding_layer.forward(b_s_)
    td_error = ddpg_for_pg_model.learn_c(es, b_a, b_r, es_, b_pg_a)
    a_loss = ddpg_for_pg_model.learn_a(es, b_pg_a)
    ddpg_for_pg_model.soft_update(ddpg_for_pg_model.Actor, ddpg_for_pg_model.Actor_)
    ddpg_for_pg_model.soft_update(ddpg_for_pg_model.Critic, ddpg_for_pg_model.Critic_)
    return y_preds, rewards, td_error, a_loss


def train(ddqn_model, ddpg_for_pg_model, model_dict, data_loader, embedding_layer, exploration_rate, device):
    """main.py:274-319: (mean critic TD error, mean summed reward per batch, epoch AUC)."""
    from sklearn.metrics import roc_auc_score
    total_loss, total_rewards, intervals = 0.0, 0.0, 0
    targets, predicts = [], []
    for features, labels in data_loader:
        features, labels = features.long().to(device), torch.unsqueeze(labels, 1).to(device)
        y_preds, rewards, td_error, _ = train_step(ddqn_model, ddpg_for_pg_model, model_dict, features, labels,
                                                   embedding_layer, exploration_rate, device)
        targets.append(labels)
        predicts.append(y_preds)
        total_loss += td_error
        total_rewards += torch.sum(rewards, dim=0).item()
        intervals += 1
    t = torch.cat(targets).cpu().numpy()
    p = torch.cat(predicts).cpu().numpy()
    return total_loss / intervals, total_rewards / intervals, roc_auc_score(t, p)


def test(ddqn_model, ddpg_for_pg_model, model_dict, embedding_layer, data_loader, loss, device):
    """main.py:321-347: (AUC, mean per-batch loss) with the greedy actions."""
    from sklearn.metrics import roc_auc_score
    targets, predicts, losses = [], [], []
    with torch.no_grad():
        for features, labels in data_loader:
            features, labels = features.long().to(device), torch.unsqueeze(labels, 1).to(device)
            ev = embedding_layer.forward(features)
            actions = ddqn_model.choose_best_action(ev)
            _, prob_weights = ddpg_for_pg_model.choose_best_action(ev, actions.float())
            y, _, _ = generate_preds(model_dict, features, actions, prob_weights, labels, device, mode="test")
            losses.append(loss(y, labels.float()).item())
            targets.append(labels)
            predicts.append(y)
    t = torch.cat(targets).cpu().numpy()
    p = torch.cat(predicts).cpu().numpy()
    return roc_auc_score(t, p), sum(losses) / len(losses)


# ---------------------------------------------------------------------------------------------------
# Data-parallel policy nets (no counterpart in the reference, which is single-device):
# every rank runs the step on its slice of the batch, the frozen tables are replicated, and the learn steps
# behave like the single-process step on the concatenated replay batch: BatchNorm batch statistics are taken
# over all ranks' samples and the parameter gradients are averaged before the optimizer step.
# ---------------------------------------------------------------------------------------------------
class _CrossRankBatchNorm(torch.autograd.Function):
    """BatchNorm1d in train mode over the samples of ALL ranks: one all_reduce of (sum, sum of squares, count) forward,
    one of (sum dy, sum dy * xhat) backward.  Each rank backpropagates its own local loss; the parameter gradients are
    summed / averaged afterwards by ``grad_sync`` like every other parameter's."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group):
        import torch.distributed as dist
        C_ = x.shape[1]
        stats = torch.cat([x.sum(0), (x * x).sum(0), x.new_tensor([float(x.shape[0])])])
        dist.all_reduce(stats, group=group)
        n = stats[-1]
        mean = stats[:C_] / n
        var = (stats[C_:2 * C_] / n - mean * mean).clamp_min_(0.0)          # biased, as BatchNorm normalises with
        invstd = torch.rsqrt(var + eps)
        xhat = (x - mean) * invstd
        if running_mean is not None:
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(mean, alpha=momentum)
                running_var.mul_(1 - momentum).add_(var * (n / (n - 1).clamp_min(1.0)), alpha=momentum)   # unbiased, as torch
        ctx.save_for_backward(xhat, invstd, weight)
        ctx.n, ctx.group = n, group
        return xhat * weight + bias

    @staticmethod
    def backward(ctx, dy):
        import torch.distributed as dist
        xhat, invstd, weight = ctx.saved_tensors
        C_ = dy.shape[1]
        dbias, dweight = dy.sum(0), (dy * xhat).sum(0)
        s = torch.cat([dbias, dweight])
        dist.all_reduce(s, group=ctx.group)
        dx = weight * invstd * (dy - s[:C_] / ctx.n - xhat * (s[C_:] / ctx.n))
        return dx, dweight, dbias, None, None, None, None, None


class CrossRankBatchNorm1d(torch.nn.BatchNorm1d):
    """``nn.BatchNorm1d`` whose train-mode statistics span the process group (same parameters, buffers and state_dict)."""
    _group = None

    def forward(self, x):
        import torch.distributed as dist
        if not self.training or not dist.is_initialized() or dist.get_world_size(self._group) == 1:
            return super().forward(x)
        if self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        return _CrossRankBatchNorm.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps,
                                         self.momentum if self.momentum is not None else 0.1, self._group)


def make_data_parallel(ddqn_model, ddpg_for_pg_model, group=None):
    """Turn the two agents of ``get_model`` into data-parallel ones over ``group`` (NCCL over NVLink on the B200 box, gloo in
    the CPU tests): parameters and buffers are broadcast from rank 0, every BatchNorm1d takes cross-rank batch statistics,
    and ``learn`` / ``learn_c`` / ``learn_a`` average the gradients over the ranks before the Adam step (one all_reduce of a
    flat buffer per network)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nets = [ddqn_model.eval_net, ddqn_model.target_net, ddpg_for_pg_model.Actor, ddpg_for_pg_model.Critic,
            ddpg_for_pg_model.Actor_, ddpg_for_pg_model.Critic_]
    for net in nets:
        for t in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(t.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        for m in net.modules():
            if type(m) is torch.nn.BatchNorm1d:
                m.__class__ = CrossRankBatchNorm1d
                m._group = group

    def grad_sync(params):
        ps = [p for p in params if p.grad is not None]
        if not ps:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        dist.all_reduce(flat, group=group)
        flat.div_(world)
        o = 0
        for p in ps:
            p.grad.copy_(flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    ddqn_model.grad_sync = grad_sync
    ddpg_for_pg_model.grad_sync = grad_sync
    return grad_sync


def make_gathered_replay(ddqn_model, ddpg_for_pg_model, group=None):
    """The cheaper data-parallel form for replay batches of a few hundred transitions: every rank keeps its own replay
    memory (the transitions of its slice of the batch), draws ``batch_size / G`` of them, the ranks' draws are joined with ONE
    all_gather per agent, and every rank runs the identical learn step on the joined batch -- the single-process step on the
    concatenated replay batch, BatchNorm statistics included, with no gradient all-reduce and no per-layer collective
    (the kernels are deterministic, so the replicas stay bit-identical).  ``make_data_parallel`` is the form for batches too
    large to replicate."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    for agent in (ddqn_model, ddpg_for_pg_model):
        if agent.batch_size % world != 0:
            raise ValueError(f"batch_size {agent.batch_size} is not a multiple of the {world} ranks")
        agent.batch_size //= world
    for net in (ddqn_model.eval_net, ddqn_model.target_net, ddpg_for_pg_model.Actor, ddpg_for_pg_model.Critic,
                ddpg_for_pg_model.Actor_, ddpg_for_pg_model.Critic_):
        for t in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(t.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)

    def gather(*tensors):
        widths = [t.shape[1] for t in tensors]
        packed = torch.cat([t.double() for t in tensors], dim=1).contiguous()      # ids < 2^53 stay exact
        out = torch.empty(world * packed.shape[0], packed.shape[1], dtype=torch.float64, device=packed.device)
        dist.all_gather_into_tensor(out, packed, group=group)
        res, o = [], 0
        for t, w in zip(tensors, widths):
            res.append(out[:, o:o + w].to(t.dtype))
            o += w
        return res

    ddqn_model.batch_gather = gather
    return gather
