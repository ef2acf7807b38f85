"""Small random PAF/BED generators for parity tests (seeded, pure Python)."""
import random

OPS_EQX = ["=", "=", "=", "X", "I", "D"]
OPS_ALL = ["=", "X", "M", "I", "D", "N", "P", "=", "I", "D"]
REF = set("MDN=X")
QRY = set("MIS=X")


def random_cigar(rng, n_ops, style="eqx", canonical=True, max_len=40, allow_zero=False):
    ops, prev = [], None
    alphabet = OPS_EQX if style == "eqx" else OPS_ALL
    for k in range(n_ops):
        c = rng.choice(alphabet)
        if k == 0 or k == n_ops - 1:
            c = rng.choice(["=", "X"] if style == "eqx" else ["=", "X", "M"])
        if canonical and c == prev:
            c = "X" if c == "=" else "="
        ln = rng.randint(1, max_len) if rng.random() < 0.25 else rng.randint(1, 4)
        if allow_zero and rng.random() < 0.1:
            ln = 0
        ops.append((ln, c))
        prev = c
    return ops


def cigar_str(ops):
    return "".join(f"{l}{c}" for l, c in ops)


def spans(ops):
    t = sum(l for l, c in ops if c in REF)
    q = sum(l for l, c in ops if c in QRY)
    return t, q


def random_paf(seed, n_contigs=3, recs_per_contig=6, style="eqx", canonical=True, allow_zero=False, lead_trail=True,
               max_ops=60, clips=False):
    """Returns (paf_text, contig_lengths).  Records on a contig overlap each other freely."""
    rng = random.Random(seed)
    lines, contigs = [], {}
    order = []
    for c in range(n_contigs):
        name = f"chr{c + 1}"
        contigs[name] = 0
        for r in range(recs_per_contig):
            order.append((name, r))
    rng.shuffle(order)  # contigs interleave in the file: emission order = first appearance
    for name, r in order:
        body = random_cigar(rng, rng.randint(1, max_ops), style, canonical, allow_zero=allow_zero)
        if lead_trail and rng.random() < 0.3:
            body = [(rng.randint(1, 4), "I") for _ in range(rng.randint(1, 2))] + body
        if lead_trail and rng.random() < 0.3:
            body = body + [(rng.randint(1, 4), rng.choice("ID")) for _ in range(rng.randint(1, 3))]
        if clips and rng.random() < 0.3:
            body = [(rng.randint(1, 5), "S")] + body
        if clips and rng.random() < 0.3:
            body = body + [(rng.randint(1, 5), "S")]
        if clips and rng.random() < 0.2:
            body = [(rng.randint(1, 5), "H")] + body
        t, q = spans(body)
        t_st = rng.randint(1, 400)
        q_st = rng.randint(0, 300)
        strand = rng.choice("+-")
        t_len = max(contigs[name], t_st + t + rng.randint(0, 50))
        contigs[name] = t_len
        q_len = q_st + q + rng.randint(0, 30)
        lines.append([f"q{len(lines)}", q_len, q_st, q_st + q, strand, name, None, t_st, t_st + t, 0, 0, rng.randint(0, 60),
                      "tp:A:P", f"cg:Z:{cigar_str(body)}", "zd:i:7"])
    for ln in lines:
        ln[6] = contigs[ln[5]]
    text = "".join("\t".join(str(x) for x in ln) + "\n" for ln in lines)
    return text.encode(), contigs


def tiling_bed(contigs, width, with_ids=False, extra_contig=True):
    rows = []
    for name, ln in contigs.items():
        st = 0
        while st < ln:
            en = min(st + width, ln)
            rows.append((name, st, en))
            st = en
    if extra_contig:
        rows.append(("chrUn_absent", 0, 100))
    return bed_text(rows, with_ids)


def random_bed(seed, contigs, n_rows, max_w=200, with_ids=True, sort=False):
    rng = random.Random(seed)
    rows = []
    names = list(contigs)
    for _ in range(n_rows):
        name = rng.choice(names)
        st = rng.randint(0, contigs[name])
        en = st + rng.randint(1, max_w)
        rows.append((name, st, en))
    if n_rows > 3:
        rows.append(rows[1])  # duplicate row -> duplicate output (Q5)
    if sort:
        rows.sort()
    return bed_text(rows, with_ids)


def bed_text(rows, with_ids):
    out = ["#comment line\n"]
    for i, (n, s, e) in enumerate(rows):
        out.append(f"{n}\t{s}\t{e}\tw{i}\n" if with_ids else f"{n}\t{s}\t{e}\n")
    return "".join(out).encode()


def random_trim_paf(seed, n_names=4, recs_per_name=4, max_ops=60, style="eqx", canonical=True, allow_zero=False, lead_trail=True,
                    span=80, big=0):
    """PAF text for `rb trim-paf`: records that share query names and overlap on the query (both strands), each starting and
    ending on an M/=/X op (after the strip of optional leading / trailing indels).  `big` > 0: that many records get ~20x the
    ops (several scan tiles per record, thousands of split-point candidates per pair)."""
    rng = random.Random(seed)
    lines = []
    for nm in range(n_names):
        for r in range(rng.randint(1, recs_per_name)):
            n_ops = rng.randint(1, max_ops) * (20 if len(lines) < big else 1)
            body = random_cigar(rng, n_ops + 2, style, canonical, allow_zero=allow_zero)
            if body[0][0] == 0:
                body[0] = (1, body[0][1])
            if body[-1][0] == 0:
                body[-1] = (1, body[-1][1])
            if lead_trail and rng.random() < 0.2:
                body = [(rng.randint(1, 4), "I")] + body
            if lead_trail and rng.random() < 0.2:
                body = body + [(rng.randint(1, 4), rng.choice("ID"))]
            t, q = spans(body)
            t_st = rng.randint(1, 4000)
            q_st = rng.randint(0, span)
            lines.append([f"query_{nm}", 10 ** 6, q_st, q_st + q, rng.choice("+-"), f"chr{rng.randint(1, 3)}", 10 ** 7, t_st, t_st + t, 0, 0,
                          rng.randint(0, 60), "tp:A:P", f"cg:Z:{cigar_str(body)}"])
    rng.shuffle(lines)  # names interleave in the file: the stable sort by query name has work to do
    return "".join("\t".join(str(x) for x in ln) + "\n" for ln in lines).encode()
