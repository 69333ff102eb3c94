"""Hyper-parameter table with the reference's flag names and defaults (runtime_config.py:6-34).

The reference builds these with ``tf.compat.v1.flags`` (absl); here ``FLAGS`` is a plain attribute
namespace and ``flags.DEFINE_*`` only adds a missing attribute, so scripts that define their own flags
(mwis_dqn_call.py:36-38) keep working and the duplicate-definition error of the reference's HEAD
("Config / flags") cannot occur.
"""
from __future__ import annotations


class _Flags:
    def __init__(self):
        self.model = "gcn_cheby"
        self.learning_rate = 0.001
        self.learning_decay = 1.0
        self.epochs = 201
        self.feature_size = 32
        self.hidden1 = 32
        self.diver_num = 32
        self.dropout = 0.0
        self.weight_decay = 5e-4
        self.early_stopping = 1000
        self.max_degree = 1
        self.num_layer = 20
        self.backoff_prob = 0.3
        self.diver_out = 32
        self.timeout = 300
        self.datapath = "./data/Random_Graph_Test"
        self.snr_db = 10.0
        self.training_set = "IS4SAT"
        self.greedy = 0
        self.skip = False
        self.wts_init = "random"
        self.snapshot = ""
        self.predict = "mwis"
        self.epsilon = 1.0
        self.epsilon_min = 0.001
        self.epsilon_decay = 0.985
        self.gamma = 1.0
        # not a reference flag: activation of the ONLY layer of a 1-layer GCN_DQN.  'identity' is the source at HEAD
        # (gcn/models.py:539-548); 'leaky_relu' is what the shipped *_l1_* checkpoints were trained with (their
        # model.ckpt.meta ends graphconvolution_1/LeakyRelu -> ArgMax; tests/test_meta_pin.py).  Scores differ only
        # where they are negative (by the factor 0.2).
        self.last_act = "identity"

    def update(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)
        return self

    def copy(self):
        c = _Flags()
        c.__dict__.update(self.__dict__)
        return c


class _FlagDefiner:
    """``flags.DEFINE_string(name, default, help)`` and friends; ``flags.FLAGS`` is the namespace."""

    def __init__(self, namespace):
        self.FLAGS = namespace

    def _define(self, name, default, _help=""):
        if not hasattr(self.FLAGS, name):
            setattr(self.FLAGS, name, default)

    DEFINE_string = DEFINE_float = DEFINE_integer = DEFINE_bool = DEFINE_boolean = _define


FLAGS = _Flags()
flags = _FlagDefiner(FLAGS)


def make_flags(**kwargs) -> _Flags:
    """A private copy of the defaults with overrides, e.g. make_flags(feature_size=1, hidden1=32,
    num_layer=20, diver_num=1) - the flag set of bash/test_dqn_500.sh."""
    return _Flags().update(**kwargs)
