"""consul_b200 — B200-native SWIM/Serf gossip simulator behind Consul's serf surface.

Only the gossip hot path of hashicorp/consul is implemented: memberlist's
failure detector + dissemination and serf's Lamport-clocked piggyback, as sm_100a CUDA
kernels behind the C ABI in include/gsim.h.
"""
from .pool import (Pool, GsimError, lan_config, wan_config, consul_test_config,  # noqa: F401
                   PRED_RUMOR_CONVERGED, PRED_ALL_RUMORS_CONVERGED, PRED_CRASHED_ALL_DEAD, NEVER,
                   FLAG_LOG_GLOBAL_EVENTS, FLAG_NO_GRAPH, FLAG_PUSH_PULL, FLAG_COORDINATES)
from .wan import WanFederation, c5_latency_matrix  # noqa: F401
