"""Ingest side of the hot path: quarterly statement JSON -> retrieval chunks -> collection.

Behavioural mirror of the reference's chunk builder (`chunking_storing (1).py:31-332`) and of its driver loop
(`:335-396`): each quarter of `extract_data/icici_q{1..4}_2023/*.json` yields up to four text chunks
(profitability, balance sheet, key ratios, segments) with the id / period / chunk_type / statement_type /
primary_value columns of the `fin_chunks` schema (`:14-22`).  The texts must equal the reference's
`FinRag_knowledge_graph/chunks.json` byte for byte - `tests/test_chunker_cpu.py` checks all 16 against the
committed copy under `tests/golden/`.

Built differently from the reference (one long function appending to a string): statements are wrapped in a small
reader, every chunk is a function returning lines, and formatting goes through three helpers.  Nothing here touches
the GPU; `ingest_chunks` hands the embeddings to the collection shim (`insert / flush / load` -> K1 on the device).
"""
from __future__ import annotations

import glob
import json
import os
from typing import Callable, Dict, Iterable, List, Optional, Sequence

# quarter directory -> period label, in the reference's processing order (`chunking_storing (1).py:339-345`)
QUARTERS = (("q1_2023", "Q1_FY2024"), ("q2_2023", "Q2_FY2024"), ("q3_2023", "Q3_FY2024"), ("q4_2023", "Q4_FY2024"))

# quarter -> (closing month, calendar-year offset of that month relative to the fiscal-year label)
_CLOSING = {"Q1": ("june", -1), "Q2": ("september", -1), "Q3": ("december", -1), "Q4": ("march", 0)}

SEGMENTS = (("retailBanking", "Retail Banking"), ("wholesaleBanking", "Wholesale Banking"), ("treasury", "Treasury"),
            ("lifeInsurance", "Life Insurance"), ("others", "Others"))


def period_columns(period: str):
    """Column keys of the current quarter and of the same quarter a year earlier, e.g. Q1_FY2024 -> june2023, june2022.
    Periods that do not mention 2024 are read as FY2023, as the reference does (`:78-90`)."""
    for tag, (month, offset) in _CLOSING.items():
        if tag in period:
            fy = 2024 if "2024" in period else 2023
            year = fy + offset
            return f"{month}{year}", f"{month}{year - 1}"
    return None, None


def _crore(x) -> str:
    return f"₹{x:,.0f} crore"


def _share(part, whole) -> float:
    return part / whole * 100 if whole else 0


def _growth(now, before) -> float:
    return (now - before) / before * 100 if before else 0


class _Line:
    """One row of a statement: a dict of column -> value; missing columns read as 0."""

    def __init__(self, node):
        self.node = node or {}

    def at(self, column):
        return self.node.get(column, 0)


class QuarterStatements:
    """Sorts the JSON documents of one quarter into the three consolidated statements the chunks are built from
    (classification rules of `chunking_storing (1).py:51-64`: report type first, then marker keys; a bare
    `segmentalResults` document is only a fallback)."""

    def __init__(self, documents: Iterable[dict]):
        self.financials = self.segments = self.balance = None
        for doc in documents:
            kind = doc.get("reportType")
            if kind == "CONSOLIDATED FINANCIAL RESULTS":
                self.financials = doc
            elif kind == "CONSOLIDATED SEGMENTAL RESULTS" or "consolidatedSegmentalResults" in doc:
                self.segments = doc
            elif "consolidatedBalanceSheet" in doc:
                self.balance = doc
            elif "segmentalResults" in doc and not self.segments:
                self.segments = doc

    @property
    def company(self) -> str:
        return self.financials.get("company", "ICICI Bank Limited")


def _chunk(period: str, suffix: str, chunk_type: str, lines: Sequence[str], primary_value) -> dict:
    return {"id": f"icici_{period.lower()}_{suffix}", "text": "".join(lines), "period": period, "chunk_type": chunk_type,
            "statement_type": "consolidated", "primary_value": primary_value}


def profitability_chunk(st: QuarterStatements, period: str, cur: str, prev: str) -> Optional[dict]:
    res = st.financials.get("consolidatedResults")
    if not res or not cur or not all(k in res for k in ("income", "expenses", "profitAndLoss")):
        return None
    inc, exp, pnl = res["income"], res["expenses"], res["profitAndLoss"]
    total_inc, prev_inc = _Line(inc["totalIncome"]).at(cur), _Line(inc["totalIncome"]).at(prev)
    interest_inc, other_inc = _Line(inc["interestEarned"]).at(cur), _Line(inc["otherIncome"]).at(cur)
    total_exp = _Line(exp["totalExpenditure"]).at(cur)
    interest_exp, operating_exp = _Line(exp["interestExpended"]).at(cur), _Line(exp["operatingExpenses"]).at(cur)
    op_profit = _Line(pnl["operatingProfit"]).at(cur)
    net, prev_net = _Line(pnl["netProfitForThePeriod"]).at(cur), _Line(pnl["netProfitForThePeriod"]).at(prev)
    provisions = _Line(pnl["provisions"]).at(cur)
    out = [f"{st.company} {period} NET PROFIT PROFITABILITY ANALYSIS:\n\n", f"NET PROFIT: {_crore(net)}"]
    if prev_net:
        out.append(f" ({_growth(net, prev_net):+.1f}% YoY growth)")
    out.append(f"\nOperating Profit: {_crore(op_profit)}")
    out.append(f"\nNet Margin: {_share(net, total_inc):.1f}% | Operating Margin: {_share(op_profit, total_inc):.1f}%\n\n")
    out.append(f"INCOME: Total {_crore(total_inc)}")
    if prev_inc:
        out.append(f" ({_growth(total_inc, prev_inc):+.1f}% YoY)")
    # the reference divides by total income unguarded here (`:136-137`): a zero total is an error there too
    out.append(f"\nInterest Income: {_crore(interest_inc)} ({interest_inc / total_inc * 100:.1f}%)")
    out.append(f"\nOther Income: {_crore(other_inc)} ({other_inc / total_inc * 100:.1f}%)\n\n")
    out.append(f"EXPENSES: Total {_crore(total_exp)}")
    out.append(f"\nInterest: {_crore(interest_exp)} | Operating: {_crore(operating_exp)}")
    out.append(f"\nProvisions: {_crore(provisions)} | Cost Ratio: {_share(total_exp, total_inc):.1f}%")
    return _chunk(period, "profitability_analysis", "profitability_analysis", out, net)


def balance_sheet_chunk(st: QuarterStatements, period: str, cur: str, prev: str) -> Optional[dict]:
    sheet = (st.balance or {}).get("consolidatedBalanceSheet")
    if not sheet or "assets" not in sheet or "capitalAndLiabilities" not in sheet:
        return None
    a, l = sheet["assets"], sheet["capitalAndLiabilities"]
    total = _Line(a["totalAssets"]).at(cur)
    advances, investments, cash = (_Line(a[k]).at(cur) for k in ("advances", "investments", "cashAndBalancesWithRBI"))
    deposits, borrowings, capital, reserves = (_Line(l[k]).at(cur) for k in ("deposits", "borrowings", "capital", "reservesAndSurplus"))
    out = [f"{st.company} {period} Balance Sheet Analysis:\n\n",
           f"ASSET COMPOSITION (Total: {_crore(total)}):\n",
           f"• Advances: {_crore(advances)} ({_share(advances, total):.1f}% of total assets)\n",
           f"• Investments: {_crore(investments)} ({_share(investments, total):.1f}% of total assets)\n",
           f"• Cash & RBI Balances: {_crore(cash)}\n\n",
           "FUNDING STRUCTURE:\n",
           f"• Customer Deposits: {_crore(deposits)}\n",
           f"• Borrowings: {_crore(borrowings)}\n",
           f"• Deposit-to-Funding Ratio: {_share(deposits, deposits + borrowings):.1f}%\n\n",
           "CAPITAL POSITION:\n",
           f"• Share Capital: {_crore(capital)}\n",
           f"• Reserves & Surplus: {_crore(reserves)}\n",
           f"• Total Equity: {_crore(capital + reserves)}"]
    return _chunk(period, "balance_sheet_health", "balance_sheet_analysis", out, total)


def ratios_chunk(st: QuarterStatements, period: str, cur: str, prev: str) -> Optional[dict]:
    res = st.financials.get("consolidatedResults")
    if not res or "ratios" not in res:
        return None
    out = [f"{st.company} {period} Key Financial Ratios & Metrics:\n\n"]
    basic = 0
    eps = res["ratios"].get("earningsPerShare")
    if eps is not None:
        basic, diluted, prev_basic = _Line(eps["basic"]).at(cur), _Line(eps["diluted"]).at(cur), _Line(eps["basic"]).at(prev)
        out += ["EARNINGS METRICS:\n", f"• Basic EPS: ₹{basic:.2f} per share"]
        if prev_basic:
            out.append(f" ({_growth(basic, prev_basic):+.1f}% YoY)")
        out.append(f"\n• Diluted EPS: ₹{diluted:.2f} per share\n\n")
    if len("".join(out)) <= 100:      # a header alone is not a chunk (`:236`)
        return None
    return _chunk(period, "key_ratios", "financial_ratios", out, basic)


def segment_chunk(st: QuarterStatements, period: str, cur: str, prev: str) -> Optional[dict]:
    doc = st.segments or {}
    body = doc.get("consolidatedSegmentalResults") or doc.get("segmentalResults")
    if not body or "segmentRevenue" not in body:
        return None
    profits = body.get("segmentResults") if "segmentResults" in body else body.get("segmentalResults")
    if profits is None:
        return None
    revenue = body["segmentRevenue"]
    rows = []
    for key, label in SEGMENTS:
        if key in revenue and cur in revenue[key]:
            r = revenue[key][cur]
            p = _Line(profits.get(key)).at(cur)
            rows.append((label, r, p, _share(p, r)))
    total = sum(r for _, r, _, _ in rows)
    rows.sort(key=lambda row: row[1], reverse=True)      # largest revenue first; equal revenues keep declaration order
    out = [f"{st.company} {period} Retail Banking & Business Segment Performance:\n\n"]
    for label, r, p, margin in rows:
        out += [f"{label.upper()} SEGMENT:\n", f"• Revenue: {_crore(r)} ({_share(r, total):.1f}%)\n",
                f"• Segment Result: {_crore(p)}\n", f"• Margin: {margin:.1f}%\n\n"]
    out.append(f"TOTAL SEGMENT REVENUE: {_crore(total)}")
    return _chunk(period, "segment_performance", "segment_analysis", out, total)


BUILDERS: Sequence[Callable] = (profitability_chunk, balance_sheet_chunk, ratios_chunk, segment_chunk)


def build_chunks(documents: Iterable[dict], period: str) -> List[dict]:
    """Chunks of one quarter, in the reference's order (profitability, balance sheet, ratios, segments)."""
    st = QuarterStatements(documents)
    if not st.financials:
        return []
    cur, prev = period_columns(period)
    chunks = []
    for build in BUILDERS:
        c = build(st, period, cur, prev)
        if c is not None:
            chunks.append(c)
    return chunks


def load_quarter(folder: str) -> List[dict]:
    docs = []
    for path in sorted(glob.glob(os.path.join(folder, "*.json"))):
        with open(path, "r") as f:
            docs.append(json.load(f))
    return docs


def build_corpus(data_folder: str) -> List[dict]:
    """All chunks of `data_folder/icici_q{1..4}_2023`, quarter by quarter (insertion order = row ids of the collection)."""
    out: List[dict] = []
    for quarter, period in QUARTERS:
        folder = os.path.join(data_folder, f"icici_{quarter}")
        if os.path.isdir(folder):
            out.extend(build_chunks(load_quarter(folder), period))
    return out


def build_corpus_from_bundle(bundle: Dict[str, Dict[str, dict]]) -> List[dict]:
    """Same as build_corpus for statements already in memory: {quarter directory: {file name: document}}
    (the layout of tests/golden/fin_statements.json)."""
    out: List[dict] = []
    for quarter, period in QUARTERS:
        docs = bundle.get(f"icici_{quarter}")
        if docs:
            out.extend(build_chunks([docs[name] for name in sorted(docs)], period))
    return out


FIELD_ORDER = ("id", "text", "embedding", "period", "chunk_type", "statement_type", "primary_value")


def ingest_chunks(collection, chunks: Sequence[dict], encode: Callable[[List[str]], "object"]):
    """`texts -> encode -> insert -> flush -> load` (`chunking_storing (1).py:379-396`).  `encode` returns one embedding
    per text: list of lists / numpy / a CPU torch tensor (host feed, as the reference's `model.encode(texts).tolist()`), or
    a CUDA torch tensor - embeddings produced on the GPU stay there and reach K1 through `ragfin_add(src_is_device=1)`
    without a host round trip.  The collection normalises and casts on the device (K1).  Returns the insert result."""
    import numpy as np
    texts = [c["text"] for c in chunks]
    emb = encode(texts)
    if not (hasattr(emb, "is_cuda") and emb.is_cuda):
        if hasattr(emb, "detach"):
            emb = emb.detach().numpy()
        emb = np.asarray(emb, dtype=np.float32)
    columns: Dict[str, list] = {name: [c[name] for c in chunks] for name in FIELD_ORDER if name != "embedding"}
    data = [emb if name == "embedding" else columns[name] for name in FIELD_ORDER]
    res = collection.insert(data)
    collection.flush()
    collection.load()
    return res
