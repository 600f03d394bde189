"""Host mirror of the v10 hybrid-TD3 main (``src/all_main/hybrid_td3_main_per_v10.py``): ``generate_preds`` (:54-164), the body of
its training loop for one batch (:354-410) and of ``test`` (:184-193).  File reading, CSV output and the epoch loop are the
caller's (the data path only); the agent is ``v10_Hybrid_TD3_model_PER.Hybrid_TD3_Model``.

Everything stays on the device: the encoder (``Feature_Embedding`` -> rlctr_featemb_fwd), the actor (3xTF32 GEMMs), the M frozen
CTR models, the ensemble scoring (rlctr_generate_preds_v10, three launches instead of the reference's O(M^2) masked ``nonzero``
passes with their host synchronisations) and the replay store.
"""
from __future__ import annotations

import torch

from .ensemble import generate_preds_v10 as generate_preds  # noqa: F401  (same name, arguments and returns as :54-55,164)

RANDOM_STEPS = 1000          # :356: the first 1000 batches act at random and only fill the replay memory


def train_step(rl_model, model_dict, features, labels, embedding_layer, batch_index, device, learn=None):
    """One batch of the loop at :354-410.  ``batch_index`` = ``i // batch_size`` of the reference (random actions and no learning
    while it is < 1000); ``learn`` overrides that switch.  Returns (y_preds, rewards, critic_loss or None)."""
    embedding_vectors = embedding_layer.forward(features)
    random = batch_index < RANDOM_STEPS
    c_actions, ensemble_c_actions, d_q_values, ensemble_d_actions = rl_model.choose_action(embedding_vectors, random)
    y_preds, rewards, return_c_actions = generate_preds(model_dict, features, ensemble_d_actions, ensemble_c_actions, c_actions,
                                                        labels, device, mode="train")
    transitions = torch.cat([features.float(), return_c_actions, d_q_values, ensemble_d_actions.float(), rewards], dim=1)   # :366-367
    rl_model.store_transition(transitions)
    critic_loss = None
    if (not random) if learn is None else learn:
        critic_loss = rl_model.learn(embedding_layer)                            # :409
    return y_preds, rewards, critic_loss


def test_batch(rl_model, model_dict, features, labels, embedding_layer, device):
    """:184-193 for one batch: (y, rewards, actions, prob_weights)."""
    with torch.no_grad():
        embedding_vectors = embedding_layer.forward(features)
        actions, c_actions, prob_weights = rl_model.choose_best_action(embedding_vectors)
        y, rewards, _ = generate_preds(model_dict, features, actions, prob_weights, c_actions, labels, device, mode="test")
    return y, rewards, actions, prob_weights
