#!/usr/bin/env python3
"""Box-scene training entrypoint - mirror of the reference's motion/crowd_ppo/main_ppo_box.py: same flags as main_ppo
with its different defaults (--test-num 10, --logdir ./log/log_box, --save-interval 1), the box-scene config
(MPVAEPolicy_samp_collision_2.yaml) and the crowd_env_2f_box.CrowdEnv semantics (2-D walkability-map penetration that always
terminates the episode)."""
import sys

from . import main_ppo as _m


def get_args(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    a = _m.get_args(argv)
    if "--test-num" not in argv:
        a.test_num = 10
    if "--logdir" not in argv:
        a.logdir = "./log/log_box"
    if "--save-interval" not in argv:
        a.save_interval = 1
    a.box_mode = True
    return a


if __name__ == "__main__":
    _m.main(get_args())
