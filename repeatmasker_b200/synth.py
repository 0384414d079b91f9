"""Synthetic workloads of SURVEY.md 8(d): a consensus library and a genome with implanted diverged copies.

Deterministic for a given seed (numpy.random.default_rng).  Used by bench.py and the parity tests; nothing here
touches the reference tree.
"""
import numpy as np

CLASSES = ["SINE/Alu", "LINE/L1", "LTR/ERV1", "DNA/hAT-Charlie", "LINE/L2", "SINE/MIR", "LTR/ERVL-MaLR", "DNA/TcMar-Tigger"]
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_TRANSITION = {ord("A"): ord("G"), ord("G"): ord("A"), ord("C"): ord("T"), ord("T"): ord("C")}
_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b


def random_dna(rng, n, gc=0.41):
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    return _ACGT[rng.choice(4, size=n, p=p)]


def make_library(seed, n_families=500, gc=0.41, mu=np.log(800.0), sigma=0.9, lo=80, hi=8000):
    """n_families consensi, lengths log-normal(mu, sigma) clipped to [lo, hi], i.i.d. composition."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.exp(rng.normal(mu, sigma, n_families)).astype(int), lo, hi)
    names = [f"fam{i:04d}#{CLASSES[i % len(CLASSES)]}" for i in range(n_families)]
    seqs = [random_dna(rng, int(n), gc).tobytes().decode() for n in lens]
    return names, seqs


def mutate(rng, seq, div, indel_rate):
    """Substitutions at rate div with ti:tv 2:1; indels at indel_rate with geometric(0.5) lengths."""
    s = np.frombuffer(seq.encode(), dtype=np.uint8).copy()
    n = len(s)
    sub = rng.random(n) < div
    ti = rng.random(n) < (2.0 / 3.0)
    for i in np.nonzero(sub)[0]:
        c = int(s[i])
        if ti[i]:
            s[i] = _TRANSITION[c]
        else:
            opts = [x for x in b"ACGT" if x != c and x != _TRANSITION[c]]
            s[i] = opts[int(rng.integers(2))]
    out = []
    i = 0
    ev = rng.random(n) < indel_rate
    while i < n:
        if ev[i]:
            k = int(rng.geometric(0.5))
            if rng.random() < 0.5:
                i += k  # deletion
                continue
            out.append(random_dna(rng, k).tobytes())  # insertion
        out.append(bytes([s[i]]))
        i += 1
    return b"".join(out).decode()


def revcomp(s):
    return _COMP[np.frombuffer(s.encode(), dtype=np.uint8)][::-1].tobytes().decode()


def make_genome(seed, length, lib_seqs, coverage=0.45, gc=0.41, div_lo=0.10, div_hi=0.30, zipf_a=1.1):
    """i.i.d. background with implanted copies: family ~ Zipf(zipf_a), strand Bernoulli(0.5), divergence
    U[div_lo, div_hi], indel rate = divergence / 8, 5' truncation to U[0.2, 1.0] of the consensus length."""
    rng = np.random.default_rng(seed)
    g = random_dna(rng, length, gc).copy()
    nf = len(lib_seqs)
    w = 1.0 / np.arange(1, nf + 1) ** zipf_a
    w /= w.sum()
    perm = rng.permutation(nf)
    covered = 0
    truth = []
    while covered < coverage * length:
        fam = int(perm[rng.choice(nf, p=w)])
        cons = lib_seqs[fam]
        keep = max(30, int(len(cons) * rng.uniform(0.2, 1.0)))
        frag = cons[len(cons) - keep:]
        d = rng.uniform(div_lo, div_hi)
        copy = mutate(rng, frag, d, d / 8.0)
        strand = int(rng.random() < 0.5)
        if strand:
            copy = revcomp(copy)
        if len(copy) >= length:
            continue
        pos = int(rng.integers(0, length - len(copy)))
        g[pos:pos + len(copy)] = np.frombuffer(copy.encode(), dtype=np.uint8)
        covered += len(copy)
        truth.append((pos, len(copy), fam, strand, d))
    return g.tobytes().decode(), truth


def fragment(seq, name, frag=60000, overlap=2000):
    """The fragments SimpleBatcher cuts a sequence longer than `frag` into, with its own arithmetic
    (SimpleBatcher.pm:642-720): the smallest divisor >= 2 with (n + (divisor-1)*overlap) / divisor <= frag, pieces of
    size = int((n + (divisor-1)*overlap) / divisor) starting every size - overlap bases, the last piece taking the
    remainder and ending at the end of the sequence; ids `<name>frag-<k>`."""
    n = len(seq)
    if n <= frag:
        return [name], [seq]
    divisor = 2
    while (n + (divisor - 1) * overlap) / divisor > frag:
        divisor += 1
    size = (n + (divisor - 1) * overlap) // divisor
    names, seqs = [], []
    for i in range(divisor - 1):
        a = i * (size - overlap)
        names.append(f"{name}frag-{i + 1}")
        seqs.append(seq[a:a + size])
    last = size + (n + (divisor - 1) * overlap) % divisor
    names.append(f"{name}frag-{divisor}")
    seqs.append(seq[n - last:])
    return names, seqs


def fragment_offsets(n, frag=60000, overlap=2000):
    """0-based start of every fragment `fragment()` cuts a sequence of n bases into."""
    if n <= frag:
        return [0]
    divisor = 2
    while (n + (divisor - 1) * overlap) / divisor > frag:
        divisor += 1
    size = (n + (divisor - 1) * overlap) // divisor
    last = size + (n + (divisor - 1) * overlap) % divisor
    return [i * (size - overlap) for i in range(divisor - 1)] + [n - last]


def write_fasta(path, names, seqs, width=50):
    with open(path, "w") as f:
        for n, s in zip(names, seqs):
            f.write(f">{n}\n")
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + "\n")
