"""Run-time configuration of the B200 engine.

The reference's config.py (:18-50) discovers a YAML file with AWS names (bucket, IAM role, Redis
instance type ...).  None of that exists on a single B200 box; what remains configurable is the
execution engine, read from the environment so that ``default()`` keeps its zero-argument form.
"""
import os


def default():
    return {
        "engine": {
            "streams": int(os.environ.get("NPW_B200_STREAMS", 4)),
            "high_priority_streams": int(os.environ.get("NPW_B200_HIGH_STREAMS", 2)),
            "inplace": os.environ.get("NPW_B200_INPLACE", "1") != "0",
        },
        "store": {"bucket": os.environ.get("NPW_B200_BUCKET", "hbm")},
    }
