"""argus flow CSV → Bot-IoT-style features (reference ``pcap_processing/feature_generator.py``).

Produces, without the dataset's 100-connection sliding window (like the reference, ``:4-5``), the
14 aggregate features of the Bot-IoT paper and optionally keeps only the 10 model features
(``--extract``).  Input: ``ra -L0 -c , -s +ltime +min +max +seq +mean +stddev +sum +spkts +sbytes
+dbytes +rate +srate +drate +dur -r file.argus > file.csv``.

Implementation note: all features are vectorised ``groupby().transform`` expressions (the
reference loops over IPs for the inbound-connection counts, ``:101-128``).
"""
from __future__ import annotations

import argparse

import pandas as pd

RENAME = {"srcaddr": "saddr", "dstaddr": "daddr", "srcbytes": "sbytes", "dstbytes": "dbytes", "srcpkts": "spkts",
          "dstpkts": "dpkts", "srcrate": "srate", "dstrate": "drate", "totpkts": "pkts", "totbytes": "bytes"}
STATE_NUMBERS = {"RST": 1, "CON": 2, "REQ": 3, "INT": 4, "URP": 5, "FIN": 6}
MODEL_FEATURES = ["seq", "stddev", "N_IN_Conn_P_SrcIP", "min", "state_number", "mean", "N_IN_Conn_P_DstIP",
                  "drate", "srate", "max"]
INBOUND_STATES = ("REQ", "CON", "EST")


def _rate(df: pd.DataFrame, keys) -> pd.Series:
    g = df.groupby(keys)
    return g["pkts"].transform("sum") / g["dur"].transform("sum")


def add_features(df: pd.DataFrame) -> pd.DataFrame:
    df = df.copy()
    df.columns = [c.lower() for c in df.columns]
    missing = [c for c in RENAME if c not in df.columns]
    if missing:
        raise KeyError(f"input is missing argus columns: {missing}")
    df = df.rename(columns=RENAME)
    df["state_number"] = df["state"].map(STATE_NUMBERS).fillna(-1).astype("int64")
    for col in ("sport", "dport"):  # ARP etc. have no ports
        df[col] = df[col].where(df[col].notna(), -1)
    df["TnBPSrcIP"] = df.groupby("saddr")["sbytes"].transform("sum")
    df["TnBPDstIP"] = df.groupby("daddr")["dbytes"].transform("sum")
    df["TnP_PSrcIP"] = df.groupby("saddr")["spkts"].transform("sum")
    df["TnP_PDstIP"] = df.groupby("daddr")["dpkts"].transform("sum")
    df["TnP_PerProto"] = df.groupby("proto")["pkts"].transform("sum")
    df["TnP_PerDport"] = df.groupby("dport")["pkts"].transform("sum")
    df["AR_P_Proto_P_SrcIP"] = _rate(df, ["saddr", "proto"])
    df["AR_P_Proto_P_DstIP"] = _rate(df, ["daddr", "proto"])
    df["AR_P_Proto_P_Sport"] = _rate(df, ["proto", "sport"])
    df["AR_P_Proto_P_Dport"] = _rate(df, ["proto", "dport"])
    inbound = df["state"].isin(INBOUND_STATES).astype("int64")
    df["N_IN_Conn_P_SrcIP"] = inbound.groupby(df["saddr"]).transform("sum")
    df["N_IN_Conn_P_DstIP"] = inbound.groupby(df["daddr"]).transform("sum")
    df["Pkts_P_State_P_Protocol_P_DestIP"] = df.groupby(["state", "proto", "daddr"])["pkts"].transform("sum")
    df["Pkts_P_State_P_Protocol_P_SrcIP"] = df.groupby(["state", "proto", "saddr"])["pkts"].transform("sum")
    return df


def generate(input_file: str, output_file: str, extract: bool = False) -> pd.DataFrame:
    df = add_features(pd.read_csv(input_file))
    if extract:
        df = df[MODEL_FEATURES]
    df.to_csv(output_file, float_format="%.3f")
    return df


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Run features generator")
    parser.add_argument("--input", "-i", type=str, required=True, help="csv input file")
    parser.add_argument("--output", "-o", type=str, required=True, help="csv output file")
    parser.add_argument("--extract", "-e", type=bool, nargs="?", const=True, default=False, help="Extract 10 useful features")
    return parser


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    generate(args.input, args.output, args.extract)


if __name__ == "__main__":
    main()
