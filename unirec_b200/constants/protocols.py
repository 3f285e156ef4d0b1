"""Enumerations whose string values are configuration API (reference: unirec/constants/protocols.py)."""
from enum import Enum


class EvaluationProtocal(Enum):
    OneVSAll = 'one_vs_all'
    OneVSK = 'one_vs_k'
    LabelAware = 'label_aware'
    SessionAware = 'session_aware'


class DataFileFormat(Enum):
    T1 = 'user-item'
    T1_1 = 'user-item-max_len'
    T2 = 'user-item-label'
    T2_1 = 'user-item-label-session'
    T3 = 'user-item-rating'
    T4 = 'user-item_group-label_group'
    T5 = 'user-item_seq'
    T5_1 = 'user_item_seq'
    T6 = 'user-item_seq-time_seq'
    T7 = 'label-index_group-value_group'


class ColNames(Enum):
    USERID = 'user_id'
    ITEMID = 'item_id'
    ITEMID_GROUP = 'item_id_list'
    LABEL = 'label'
    LABEL_GROUP = 'label_list'
    USER_HISTORY = 'item_seq'
    TIME_HISTORY = 'time_seq'
    SESSION = 'session_id'
    MAX_LEN = 'max_len'


class HistoryMaskMode(Enum):
    Unorder = 'unorder'
    Autoregressive = 'autoregressive'


class TaskType(Enum):
    TRAIN = 'train'
    TEST = 'test'
    INFER = 'infer'
