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
    ddqn_model.learn(embedding_layer.forward(b_s), b_a, b_r, embedding_layer.forward(b_s_))
    b_s, b_a, b_r, b_s_, b_pg_a = ddpg_for_pg_model.sample_batch()
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
