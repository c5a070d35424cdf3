"""rabi_b200 -- B200-native (sm_100a) backend for RABI's MAF hot path.

Drop-in points (hydra configs of the reference, src/rbibm/config):
    model:   module_path: rabi_b200.models   class_name: InverseAffineAutoregressiveModel
    attack:  attack_module: rabi_b200.attacks  attack_class: L2PGDAttack  loss_module: rabi_b200.loss
    metric:  metric_module: rabi_b200.metric   metric_class: ReverseKLRobMetric
    defense: defense_module: rabi_b200.defenses  defense: FIMTraceRegularizer | L2PGDrKLTrades
"""
__version__ = "0.1.0"
