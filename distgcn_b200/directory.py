"""Model / result folder naming of the reference (directory.py:5-40)."""
from __future__ import annotations

import os


def find_model_folder(FLAGS, postfix):
    """./model/result_{training_set}_deep_ld{feature_size}_c{hidden1}_l{num_layer}_cheb{max_degree}_
    diver{diver_num}_{predict}_{postfix}[/snapshot]   (directory.py:33-40)."""
    model_origin = "result_{}_deep_ld{}_c{}_l{}_cheb{}_diver{}_{}_{}".format(
        FLAGS.training_set, FLAGS.feature_size, FLAGS.hidden1, FLAGS.num_layer, FLAGS.max_degree, FLAGS.diver_num,
        FLAGS.predict, postfix)
    model_origin = os.path.join("./model", model_origin)
    if hasattr(FLAGS, "snapshot"):
        if not FLAGS.snapshot == "":
            model_origin = os.path.join(model_origin, FLAGS.snapshot)
    return model_origin


def create_result_folder(FLAGS, postfix):
    """Result folder name of directory.py:5-30 (created on demand)."""
    data_path = FLAGS.datapath
    if FLAGS.greedy == 1:
        greedy_string = "_greedy"
    elif FLAGS.greedy == 2:
        greedy_string = "_greedy_snr{}".format(FLAGS.snr_db)
    else:
        greedy_string = "_" + FLAGS.predict
    initstr = "zeros" if FLAGS.wts_init == "zeros" else ""
    skipstr = "_skip" if FLAGS.skip else "_no_skip"
    outputfolder = "./res_{:04d}_{}_{}_{}_{}_{}{}{}_{}".format(
        FLAGS.timeout, FLAGS.training_set + initstr, FLAGS.diver_num, FLAGS.diver_out, FLAGS.backoff_prob,
        data_path.split("/")[-1], greedy_string, skipstr, postfix)
    if not os.path.isdir(outputfolder):
        os.makedirs(outputfolder)
    return outputfolder
