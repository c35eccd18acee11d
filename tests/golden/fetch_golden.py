#!/usr/bin/env python3
"""Copies the reference's known-answer fixtures for the hot path into tests/golden/ (verbatim, no edits):

  models/deer ink studios/*.json   six bundled LSTM-12 captures, each with 2048-sample input_batch /
                                   output_batch vectors that the reference itself checks at 1e-5
                                   (rt-neural-generic.cpp:900-955).  (c) deer ink studios, CC BY-NC-ND 4.0:
                                   redistributed unmodified, with attribution, for testing only.
  modules/RTNeural/models/{lstm_1d,gru_1d,bad_lstm}.json + test_data/{lstm_1d,gru_1d}_{x,y}_python.csv
                                   RTNeural's Keras-generated layer vectors (BSD-3-Clause, (c) jatinchowdhury18).

/root/reference does not exist on the GPU box, so the fixtures are committed; this script documents
where they came from and regenerates them (python tests/golden/fetch_golden.py)."""
import glob
import os
import shutil

REF = os.environ.get("AIDAX_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    n = 0
    for p in sorted(glob.glob(os.path.join(REF, "models", "deer ink studios", "*.json"))):
        shutil.copyfile(p, os.path.join(HERE, "models", os.path.basename(p)))
        n += 1
    rt = os.path.join(REF, "modules", "RTNeural")
    for name in ("lstm_1d.json", "gru_1d.json", "bad_lstm.json"):
        shutil.copyfile(os.path.join(rt, "models", name), os.path.join(HERE, "rtneural", name))
        n += 1
    for name in ("lstm_1d_x_python.csv", "lstm_1d_y_python.csv", "gru_1d_x_python.csv", "gru_1d_y_python.csv"):
        shutil.copyfile(os.path.join(rt, "test_data", name), os.path.join(HERE, "rtneural", name))
        n += 1
    print("copied %d fixture files" % n)


if __name__ == "__main__":
    main()
