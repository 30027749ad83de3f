"""Weight-sharing types per node kind (reference: struct/NodeWeightSharingConfig.py:4-40):
0 none | 1 per edge, per iteration | 2 per node, per iteration | 3 one scalar per iteration |
4 per edge, shared over iterations | 5 per node, shared over iterations."""
from .NodeType import NodeType


class NodeWeightSharingConfig:
    def __init__(self, cn_weight_sharing: int, ucn_weight_sharing: int, vn_weight_sharing: int):
        self.cn_weight_sharing = cn_weight_sharing
        self.ucn_weight_sharing = ucn_weight_sharing
        self.vn_weight_sharing = vn_weight_sharing

    def __iter__(self):
        yield (NodeType.CN, self.cn_weight_sharing)
        yield (NodeType.UCN, self.ucn_weight_sharing)
        yield (NodeType.VN, self.vn_weight_sharing)

    def get(self, node_type: NodeType):
        return {NodeType.CN: self.cn_weight_sharing, NodeType.UCN: self.ucn_weight_sharing,
                NodeType.VN: self.vn_weight_sharing}.get(node_type)
