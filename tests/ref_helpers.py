"""Access to the UNMODIFIED reference for the tests that need it: /root/reference in the build container, the staged
copy under baseline/_ref/ (tools/stage_reference.py) on the GPU box."""
import os
import re
import sys
from unittest.mock import MagicMock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_dir():
    for d in (os.environ.get("RFMSC_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.isfile(os.path.join(d, "RFN", "RFN_new.py")) and os.path.isdir(os.path.join(d, "Flow")):
            return d
    return None


def reference_args(ref, argv=(), script="main_rfn.py"):
    """argparse Namespace of main_rfn.py (its defaults = configuration D) or another main_*.py with `argv` overrides,
    without importing the trainer (which needs matplotlib)."""
    src = open(os.path.join(ref, script)).read()
    head = src[:src.index("if __name__")]
    body = src[src.index("if __name__"):]
    body = body[body.index("\n") + 1:body.index("args = parser.parse_args()")]
    body = "\n".join(l[4:] if l.startswith("    ") else l for l in body.split("\n"))
    ns = {}
    exec(re.sub(r"^from .*$|^import (?!argparse).*$", "", head, flags=re.M) + "\nimport argparse\n" + body, ns)
    return ns["parser"].parse_args(list(argv))


JOB_SCRIPT_ARGV = ("--extractor_structure 16-16-pool-32 32-pool-64 64-pool-128 128-pool-256 256-pool-512 "
                   "--upscaler_structure 256 upsample-128-128 upsample-64-64 upsample-32-32 upsample-16-16 "
                   "--prior_structure 256 256 --encoder_structure 256 256 --make_conditional --learn_prior "
                   "--skip_connection_features --flow_norm actnorm --structure_scaler 2 --choose_data mnist "
                   "--n_units_affine 256 --n_units_prior 512 --temperature 0.7 --norm_type none --z_dim 56 --h_dim 200 "
                   "--n_bits 8 --n_frames 10 --K 10 --L 5 --skip_connection_flow without_skip --no-upscaler_tanh "
                   "--no-downscaler_tanh").split()   # RFN/default_rfn_job.sh:85 (configuration J)


def job_script_args(ref, batch, extra=()):
    argv = list(JOB_SCRIPT_ARGV) + ["--batch_size", str(batch), "--x_dim", str(batch), "1", "64", "64",
                                    "--condition_dim", str(batch), "1", "64", "64"] + list(extra)
    return reference_args(ref, argv)


def stub_optional_imports():
    for m in ["matplotlib", "matplotlib.pyplot", "imageio", "torchfile", "parse"]:
        sys.modules.setdefault(m, MagicMock())


def purge_reference_modules():
    for name in [n for n in sys.modules if n.split(".")[0] in ("Flow", "Utils", "RFN", "SRNN")]:
        del sys.modules[name]
