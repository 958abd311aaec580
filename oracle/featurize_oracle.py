"""TEST INFRASTRUCTURE -- CPU restatement of the reference's host featurisation, `Config.from_list_to_tensor`
(config/Config.py:162-233), the producer of every tensor the graph hot path receives from the host.

Only tests/ (and oracle/pin_featurize.py) may import this.  Pinned: oracle/pin_featurize.py executes the reference's
own function body (extracted from /root/reference/config/Config.py by AST, Config.py itself is not importable here)
on synthetic records and requires equality on every output; tests/golden/featurize.npz holds outputs of the reference
itself.  Plain numpy loops, small cases only.
"""
from __future__ import annotations

import numpy as np


def make_dis2idx() -> np.ndarray:
    """config/Config.py:106-116."""
    t = np.zeros(1024, dtype="int64")
    t[1] = 1
    t[2:] = 2
    t[4:] = 3
    t[8:] = 4
    t[16:] = 5
    t[32:] = 6
    t[64:] = 7
    t[128:] = 8
    t[256:] = 9
    t[512:] = 10
    return t


def from_list_to_tensor(item, max_length: int = 512, max_num: int = 5, dis_plus: int = 10) -> dict:
    """numpy outputs with the reference's dtypes (C:219-225): adj_matrix f32, sen_matrix bool, pos_matrix_h/t i64,
    node_pos f32, node_type i64, node_relative_pos i64."""
    dis2idx = make_dis2idx()
    graph = item["graph"]
    doc_len = len(item["document"])
    n = len(graph.nodes())
    node_pos = np.zeros((n, doc_len))                                        # C:169
    node_type = np.zeros(n)
    for node in graph.nodes():                                               # C:171-176
        for position in graph.nodes[node]["exist_pos"]:
            node_pos[node, position[0]:position[1]] = 1.0 / (position[1] - position[0])
        node_pos[node, :] *= 1.0 / len(graph.nodes[node]["exist_pos"])
        node_type[node] = graph.nodes[node]["type"][0]
    S = graph.graph["max_sentence_num"]
    adj = np.zeros((n, n))                                                   # C:178
    sen = np.zeros((n, n, S, doc_len))                                       # C:180-182
    pos_h = np.zeros((n, n, S, doc_len))
    pos_t = np.zeros((n, n, S, doc_len))
    for u, v, edge in graph.edges(data=True):                                # C:183-205
        adj[u, v] = 1
        for j, (sentence, position) in enumerate(zip(edge["sentences"], edge["position"])):
            sen[u, v, j, sentence[0]:sentence[1]] = 1
            for k in range(sentence[0], sentence[1]):
                dl, dr = k - position[0], k - position[1]
                if dl < 0:
                    pos_h[u, v, j, k] = -int(dis2idx[-dl])
                elif dr > 0:
                    pos_h[u, v, j, k] = int(dis2idx[dr])
                dl, dr = k - position[2], k - position[3]
                if dl < 0:
                    pos_t[u, v, j, k] = -int(dis2idx[-dl])
                elif dr > 0:
                    pos_t[u, v, j, k] = int(dis2idx[dr])
                pos_h[u, v, j, k] += dis_plus
                pos_t[u, v, j, k] += dis_plus
    rel = np.zeros((n, n))                                                   # C:207-217
    for a in graph.nodes():
        for b in graph.nodes():
            if a == b:
                continue
            d = graph.nodes[a]["exist_pos"][0][0] - graph.nodes[b]["exist_pos"][0][0]
            rel[a, b] = -dis2idx[-d] if d < 0 else dis2idx[d]
    return {                                                                 # C:219-225
        "adj_matrix": adj.astype(np.float32),
        "sen_matrix": sen[:, :, :max_num, :max_length].astype(bool),
        "pos_matrix_h": pos_h[:, :, :max_num, :max_length].astype(np.int64),
        "pos_matrix_t": pos_t[:, :, :max_num, :max_length].astype(np.int64),
        "node_pos": node_pos[:, :max_length].astype(np.float32),
        "node_type": node_type.astype(np.int64),
        "node_relative_pos": rel.astype(np.int64),
    }
