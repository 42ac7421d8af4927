"""The per-VBlock codec path over a batch of aligned, coordinate-sorted BAM VBlocks (BASELINE.json configs[2]) — the same
pipelines as fastq_path.py with the stream set the SAM/BAM segmenter leaves behind (the segmenter itself, the reference and the
aligner are out of scope; what reaches the codecs is synthesised here, post-seg, as SURVEY §8d C3 describes):

  QUAL      QUAL.local  --codec_domq_compress-->  QUAL / DOMQRUNS / QUALMPLX / DIVRQUAL --sub-codecs-->  sections   (as FASTQ)
  SEQ       the reference explains most bases: SQBITMAP.local (LT_BITMAP, one bit per base, src/sam_seq.c:189) and STRAND.local
            (LT_BITMAP, one bit per read, :190) and GPOS.local (uint32 per read) go through the simple codecs; only the bases the
            reference does not explain reach codec_acgt as NONREF.local (hard-coded CODEC_ACGT, src/sam_seg.c:820; padded to a
            multiple of four bases, src/sam_seq.c:221) — a small fraction of what a FASTQ VBlock hands over
  fields    FLAG / POS / MAPQ / CIGAR / TLEN b250 and local streams (MAPQ is LT_UINT8, src/sam_seg.c:397), the aux integers
            NM:i / AS:i / XS:i (dyn-int locals, :407-408), MD:Z, and the QNAME contexts  --simple codecs-->  sections

One class, one set of kernels: BamCodecPath is FastqCodecPath with these fields and NONREF's length.
"""
import torch

from .fastq_path import FastqCodecPath, _synth_chunk

NONREF_FRAC = 0.015          # bases handed to codec_acgt: ~0.5 % mismatches + ~1 % unmapped reads


def nonref_len(n_reads, read_len):
    return 4 * ((int(n_reads * read_len * NONREF_FRAC) + 3) // 4)


def bam_fields(n_reads, read_len):
    """{context: bytes per VBlock} of the simple-codec streams of an aligned BAM VBlock (in the order they are compressed)"""
    n = n_reads * read_len
    return {"SQBITMAP": (n + 7) // 8, "STRAND": (n_reads + 7) // 8, "GPOS": 4 * n_reads,
            "FLAG": n_reads, "POS": n_reads, "MAPQ": n_reads, "CIGAR": n_reads, "TLEN": 2 * n_reads,
            "NM_i": n_reads, "AS_i": n_reads, "XS_i": n_reads, "MD_Z": n_reads,
            "Q_TILE": n_reads, "Q_X": 4 * n_reads, "Q_Y": 4 * n_reads, "Q_MISC": n_reads}


def bam_bytes_per_vb(n_reads, read_len):
    """uncompressed BAM record bytes a VBlock of n_reads represents: block_size + the 32 fixed bytes, read name (40), one CIGAR op,
    4-bit SEQ, QUAL, and the aux fields NM:C AS:C XS:C MD:Z RG:Z (~28)"""
    return n_reads * (4 + 32 + 40 + 4 + (read_len + 1) // 2 + read_len + 28)


def synth_bam_vblocks(V, n_reads, read_len, seed, device):
    """Synthetic post-seg streams of aligned coordinate-sorted VBlocks: dict of uint8 tensors [V, ...] — qual (as FASTQ: binned
    Illumina qualities with run structure, DOMQ's case), seq = NONREF.local, and the fields of bam_fields()."""
    parts = [_bam_chunk(min(8, V - v0), n_reads, read_len, seed * 100003 + v0, device) for v0 in range(0, V, 8)]
    return {k: torch.cat([p[k] for p in parts], 0).contiguous() for k in parts[0]}


def _bam_chunk(V, n_reads, read_len, seed, device):
    d = _synth_chunk(V, n_reads, read_len, seed, device)                   # qual + the QNAME contexts (seq is replaced below)
    g = torch.Generator(device=device); g.manual_seed(seed ^ 0x5bd1e995)
    n = n_reads * read_len
    u8 = dict(dtype=torch.uint8, device=device)
    rnd = lambda *shape: torch.rand(shape, generator=g, device=device)

    def pick(probs, *shape):
        k = 1
        for s in shape: k *= s
        return torch.multinomial(torch.tensor(probs, device=device), k, replacement=True, generator=g).view(*shape)

    d["seq"] = d["seq"][:, :nonref_len(n_reads, read_len)].contiguous()     # the bases the reference does not explain: uniform ACGT, 0.1 % N
    # SQBITMAP: bit = 1 where the base equals the reference; 0.5 % mismatches, 1 % of the reads unmapped (no bits set)
    match = rnd(V, n_reads, read_len) >= 0.005
    match &= (rnd(V, n_reads, 1) >= 0.01)
    bits = match.view(V, n).to(torch.uint8)
    pad = (-n) % 8
    if pad:
        bits = torch.cat([bits, torch.zeros((V, pad), **u8)], 1)
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], **u8)
    d["SQBITMAP"] = (bits.view(V, -1, 8) * w).sum(2).to(torch.uint8).contiguous()
    sb = (rnd(V, n_reads) < 0.5).to(torch.uint8)
    pad = (-n_reads) % 8
    if pad:
        sb = torch.cat([sb, torch.zeros((V, pad), **u8)], 1)
    d["STRAND"] = (sb.view(V, -1, 8) * w).sum(2).to(torch.uint8).contiguous()
    # GPOS: genome position of each read, big-endian uint32; coordinate-sorted at 30x: ~5 bases between consecutive reads
    gpos = (torch.cumsum(torch.randint(0, 11, (V, n_reads), generator=g, device=device), 1) + 10_000_000).to(torch.int32)
    d["GPOS"] = gpos.contiguous().view(torch.uint8).view(V, n_reads, 4).flip(2).contiguous().view(V, 4 * n_reads)
    d["FLAG"] = pick([.24, .24, .24, .24, .01, .01, .01, .01], V, n_reads).to(torch.uint8).contiguous()   # b250: word indices of 99 / 147 / 83 / 163 / ...
    d["POS"] = torch.randint(0, 11, (V, n_reads), generator=g, device=device).to(torch.uint8)          # b250 of the delta snips
    d["MAPQ"] = torch.tensor([60, 0, 27, 40, 48], **u8)[pick([.9, .04, .02, .02, .02], V, n_reads)].contiguous()
    d["CIGAR"] = torch.where(rnd(V, n_reads) < 0.92, torch.zeros((), dtype=torch.long, device=device),
                             torch.randint(1, 200, (V, n_reads), generator=g, device=device)).to(torch.uint8)
    tlen = (torch.randn((V, n_reads), generator=g, device=device) * 60 + 420).clamp(150, 2000).to(torch.int16)
    d["TLEN"] = tlen.contiguous().view(torch.uint8).view(V, n_reads, 2).flip(2).contiguous().view(V, 2 * n_reads)
    nm = (~match).sum(2).clamp(max=255).to(torch.uint8)                                               # NM:i = mismatches of the read
    d["NM_i"] = nm.contiguous()
    d["AS_i"] = (read_len - 5 * nm.to(torch.int32)).clamp(0, 255).to(torch.uint8).contiguous()
    d["XS_i"] = torch.where(rnd(V, n_reads) < 0.7, torch.zeros((), dtype=torch.long, device=device),
                            torch.randint(19, read_len + 1, (V, n_reads), generator=g, device=device)).clamp(max=255).to(torch.uint8)
    d["MD_Z"] = torch.where(nm == 0, torch.zeros((), dtype=torch.long, device=device),
                            torch.randint(1, 250, (V, n_reads), generator=g, device=device)).to(torch.uint8)
    return d


class BamCodecPath(FastqCodecPath):
    """FastqCodecPath over the streams of an aligned BAM VBlock: DOMQ on every quality, codec_acgt on NONREF only, sixteen
    field streams through the simple codecs"""

    def __init__(self, eng, V, n_reads, read_len, n_engines=3, sub_batch=128):
        super().__init__(eng, V, n_reads, read_len, n_engines=n_engines, sub_batch=sub_batch,
                         fields=bam_fields(n_reads, read_len), seq_len=nonref_len(n_reads, read_len))
