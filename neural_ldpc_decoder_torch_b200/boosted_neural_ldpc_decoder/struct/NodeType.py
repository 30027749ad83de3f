from enum import Enum


class NodeType(Enum):
    CN = "CN"
    UCN = "UCN"
    VN = "VN"
