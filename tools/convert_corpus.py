#!/usr/bin/env python
"""Convert the reference's processed dataset file to the GPU loader's corpus format (analysisgnn_b200/corpusfile.py).

    python tools/convert_corpus.py processed/data.pt corpus.agc [--rel onset consecutive during rest]
    python tools/convert_corpus.py --info corpus.agc

Run the conversion where torch_geometric is installed (the file pickles a ``HeteroData`` class object:
analysisgnn/data/data_utils.py:53, 79-80); reading the result needs only numpy and torch.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from analysisgnn_b200 import corpusfile  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("src", nargs="?")
    ap.add_argument("dst", nargs="?")
    ap.add_argument("--rel", nargs="+", default=list(corpusfile.REL_NAMES), help="note->note relations, in id order")
    ap.add_argument("--info", metavar="FILE", help="print the array table of a corpus file and verify its checksums")
    args = ap.parse_args()
    if args.info:
        tab = corpusfile.read_table(args.info)
        corpusfile.load_arrays(args.info, verify=True)
        print(f"{args.info}: {tab['n_rel']} relations, checksums ok")
        for e in tab["arrays"]:
            print(f"  {e['name']:24s} {e['dtype']:8s} {str(tuple(e['shape'])):20s} offset {e['offset']:>12d}  {e['nbytes']:>12d} B")
        return
    if not args.src or not args.dst:
        ap.error("source and destination files are required")
    saved = torch.load(args.src, weights_only=False)
    data, slices = saved[0], saved[1]
    if not isinstance(data, dict):               # older PyG versions save the collated object itself
        data = data.to_dict()
    corpus = corpusfile.from_pyg_collated(data, slices, rel_names=args.rel)
    corpusfile.save_corpus(args.dst, corpus)
    print(f"{args.dst}: {corpus.n_scores} scores, {corpus.node_ptr[-1]} notes, {corpus.edges.shape[1]} edges, "
          f"extras {sorted(corpus.extras)}")


if __name__ == "__main__":
    main()
