"""Command-line drivers with the reference's flags (Anchored_Fusion.py:15-30,
Anchored_Fusion_singlecell.py:16-29), running the anchoring stage on the GPU.

Scope: these drivers do what the reference scripts do up to and including the anchoring stage
(gene-name parsing, per-gene anchor FASTA, the stage's output files) and then add the
record interpretation that sits directly on those records (split points via contact_reads).
The later stages (genome bwa, BLAT, the CNN/Transformer filter) are out of scope; because every
stage of the reference is guarded by the existence of its output file, running the reference
script afterwards on the same --out_folder picks up from here.
"""
import argparse
import os
import re
import sys
import time

from .functions import combine_split_reads, split_records
from .stage import GeneAnchorer, anchor_cells, anchor_stage, anchor_stage_multi


def _common_flags(p):
    p.add_argument('--file_anchored_cds', type=str, required=True, default='', help='Target gene fasta file of anchored transcript')
    p.add_argument('--gene_names', type=str, default='', help='The file of target gene names')
    return p


def _tail_flags(p):
    p.add_argument('--out_folder', type=str, default='./', help='The folder of the output file')
    p.add_argument('--file_ref_seq', type=str, default='', help='The reference sequence file (used by later stages only)')
    p.add_argument('--file_ref_ann', type=str, default='', help='The reference annotation file (used by later stages only)')
    p.add_argument('--not_filter_false_positive', action='store_true', help='Accepted for compatibility; the filter is a later stage.')
    p.add_argument('--not_train_filter_model', action='store_true', help='Accepted for compatibility; the filter is a later stage.')
    p.add_argument('--model_file', type=str, default='./data/model.pt', help='Accepted for compatibility.')
    p.add_argument('--positive_samples', type=str, default='./data/positive_samples.txt', help='Accepted for compatibility.')
    p.add_argument('--homo_gene_file', type=str, default='./data/homo_gene.npy', help='Accepted for compatibility.')
    p.add_argument('--negative_samples', type=str, default='./Model/negative_samples.txt', help='Accepted for compatibility.')
    p.add_argument('--thread', type=str, default='1', help='The threads number you want use (FASTQ decode / pack workers; the reference hands it to bwa mem -t).')
    p.add_argument('--gpu_number', type=str, default='-1', help="The gpu number you want use ('-1': first visible GPU).")
    return p


_NOISE = re.compile(r'gene|specie|trans|for|homo|sapiens', re.IGNORECASE)


def parse_gene_names(file_anchored_cds, gene_names_file=''):
    """Gene names as the reference derives them (Anchored_Fusion.py:58-80): from --gene_names, one per
    line, else from the FASTA headers with accession-like tokens and annotation words dropped."""
    if gene_names_file and os.path.exists(gene_names_file):
        with open(gene_names_file) as fh:
            return [l.rstrip() for l in fh if l.rstrip() != '']
    names = []
    with open(file_anchored_cds) as fh:
        for line in fh:
            if not line.startswith('>'):
                continue
            tokens = [t for t in line.rstrip()[1:].split(' ')
                      if not re.match(r'[a-zA-Z]+_\d+\.\d+', t) and not _NOISE.search(t)]
            names.append(tokens[0])
    return names


def split_anchor_fasta(file_anchored_cds, gene_names, path_of):
    """Write one <work>_anchored_gene_sequence.fa per gene (Anchored_Fusion.py:123-165).  QUIRK kept:
    the reference starts at line 1 and skips exactly one header per gene, i.e. the i-th FASTA
    record goes to the i-th gene name."""
    with open(file_anchored_cds) as fh:
        lines = fh.readlines()
    i, out = 1, []
    for gene in gene_names:
        path = path_of(gene)
        with open(path, 'w') as o:
            o.write('>' + gene + '\n')
            while i < len(lines) and not lines[i].startswith('>'):
                o.write(lines[i])
                i += 1
        i += 1
        out.append(path)
    return out


def _mkdir(p):
    os.makedirs(p, exist_ok=True)


def write_split_points(stats, gene, path):
    """type / split point / support table from the raw anchored records (contact_reads semantics,
    functions.py:892-952, before the genome-contiguity filter of del_too_many_reads)."""
    rows = []
    with open(stats['raw_sam']) as fh:
        for line in fh:
            a = line.rstrip('\n').split('\t')
            rows.append((a[0], a[2], a[3], a[5], a[9]))
    groups = combine_split_reads(split_records(rows))
    with open(path, 'w') as o:
        o.write('gene\tsplit_point\ttype\tsupport\tseq_left\tseq_right\n')
        for g in sorted(groups, key=lambda g: (-g.cnt, g.breakpoint)):
            o.write('%s\t%d\t%s\t%d\t%s\t%s\n' % (gene, g.breakpoint, g.type_, g.cnt, g.seq_left, g.seq_right))
    return groups


def contiguity_stage(out_dir_name, gene, args):
    """del_too_many_reads (Anchored_Fusion.py:202-203) when --file_ref_seq names a genome FASTA: the 2-op anchored
    reads are re-aligned to the genome on the GPU (genome.py; the reference runs `bwa mem` on the genome here) and
    the survivors go to `<w>_anchored_reads.sam` -- the file the reference's Find_fine_block / contact_reads read
    next, and whose existence makes the reference skip its own del_too_many_reads.  Also writes the split-point
    table of the survivors.  Returns the number of surviving reads, or None when there is no genome."""
    ref = getattr(args, 'file_ref_seq', '')
    if not ref or not os.path.isfile(ref):
        return None
    from .functions import contact_reads, del_too_many_reads
    out_sam = out_dir_name + '_anchored_reads.sam'
    if not os.path.exists(out_sam):
        part = out_sam + '.partial'
        from .stage import resolve_device
        del_too_many_reads(out_dir_name + '_anchored_reads.bam', part, out_dir_name, ref, args.thread, device=resolve_device(args.gpu_number))
        os.replace(part, out_sam)
    groups = contact_reads(out_sam, out_dir_name, ref, args.thread)
    with open(out_dir_name + '_split_points_filtered.txt', 'w') as o:
        o.write('gene\tsplit_point\ttype\tsupport\tseq_left\tseq_right\n')
        for g in sorted(groups, key=lambda g: (-g.cnt, g.breakpoint)):
            o.write('%s\t%d\t%s\t%d\t%s\t%s\n' % (gene, g.breakpoint, g.type_, g.cnt, g.seq_left, g.seq_right))
    with open(out_sam) as fh:
        return sum(1 for _ in fh)


class ContiguityBatcher:
    """The contiguity stage for many (cell, gene) outputs at once: the 2-op reads of many cells share the passes over the
    genome (one pass serves ~86 reads whoever they belong to; a 5 000-pair cell has a handful), the decision and the
    files stay per cell.  Same files as contiguity_stage."""

    def __init__(self, args, flush_reads=2048):
        self.args, self.flush_reads = args, flush_reads
        ref = getattr(args, 'file_ref_seq', '')
        self.enabled = bool(ref) and os.path.isfile(ref)
        self.items, self.n_reads = [], 0

    def add(self, out_dir_name, gene):
        if not self.enabled or os.path.exists(out_dir_name + '_anchored_reads.sam'):
            return
        from .bam import read_bam, sam_line
        from .functions import two_op_records
        recs = list(two_op_records([sam_line(r) for r in read_bam(out_dir_name + '_anchored_reads.bam')[2]]))
        self.items.append((out_dir_name, gene, recs))
        self.n_reads += len(recs)
        if self.n_reads >= self.flush_reads:
            self.flush()

    def flush(self):
        if not self.items:
            return
        from .functions import contact_reads, contiguity_filter
        from .genome import genome_for
        from .stage import resolve_device
        g = genome_for(self.args.file_ref_seq, resolve_device(self.args.gpu_number))
        seqs = [s for _, _, recs in self.items for _, s in recs]
        hits = g.align([s if len(s) <= 256 else '' for s in seqs]) if seqs else []
        by_read = {int(h['read_id']): h for h in hits}
        base = 0
        for out_dir_name, gene, recs in self.items:
            mine = [dict_hit(by_read[base + i], i) for i in range(len(recs)) if base + i in by_read]
            lines = g.sam_lines([t for t, _ in recs], [s for _, s in recs], mine)
            base += len(recs)
            out_sam = out_dir_name + '_anchored_reads.sam'
            with open(out_sam + '.partial', 'w') as fo:
                fo.writelines(contiguity_filter(lines))
            os.replace(out_sam + '.partial', out_sam)
            groups = contact_reads(out_sam, out_dir_name, self.args.file_ref_seq, self.args.thread)
            with open(out_dir_name + '_split_points_filtered.txt', 'w') as o:
                o.write('gene\tsplit_point\ttype\tsupport\tseq_left\tseq_right\n')
                for gr in sorted(groups, key=lambda gr: (-gr.cnt, gr.breakpoint)):
                    o.write('%s\t%d\t%s\t%d\t%s\t%s\n' % (gene, gr.breakpoint, gr.type_, gr.cnt, gr.seq_left, gr.seq_right))
        self.items, self.n_reads = [], 0


def dict_hit(h, local_id):
    """a genome record re-addressed to its cell's own read numbering (Genome.sam_lines looks records up by read_id)"""
    return {'read_id': local_id, 'pos': int(h['pos']), 'clip_l': int(h['clip_l']), 'm_len': int(h['m_len']),
            'clip_r': int(h['clip_r']), 'score_strand': int(h['score_strand'])}


def run_gene_sample(file_anchored_seq, gene, fastq1, fastq2, out_dir_name, args, gene_anchorer=None):
    done = out_dir_name + '_anchored_reads.bam'
    if os.path.exists(done) and os.path.exists(out_dir_name + '_realign_reads.bam'):
        print('[anchoring] %s: outputs exist, skipping (same existence guard as the reference)' % out_dir_name)
        return None
    t0 = time.time()
    stats = anchor_stage(file_anchored_seq, fastq1, fastq2, out_dir_name, thread=args.thread,
                         gpu_number=args.gpu_number, gene_name=gene, gene_anchorer=gene_anchorer)
    if stats is None:               # a rank other than 0 of a torchrun job: rank 0 writes the files
        return None
    try:
        groups = write_split_points(stats, gene, out_dir_name + '_split_points.txt')
    except IndexError:
        # the reference's pile-up (Co_Split_reads, functions.py:166) holds 200 bases on either side of a junction and
        # raises on a longer side (2x300 reads); the stage's own files above are complete, only this summary table is not
        print('[anchoring] %s: a split read has more than 200 bases on one side of its junction -- the reference\'s '
              'Co_Split_reads array (400 columns) cannot hold it; %s_split_points.txt not written' % (gene, out_dir_name))
        groups = []
    kept = contiguity_stage(out_dir_name, gene, args)
    if kept is not None:
        print('[anchoring] %s: %d 2-op reads survive the genome-contiguity filter' % (gene, kept))
    dt = time.time() - t0
    print('[anchoring] %s: %d pairs, %d anchored reads, %d half-anchored pairs, %d split-point groups, %.2f s (%.0f pairs/s)'
          % (gene, stats['pairs'], stats['anchored'], stats['half_anchored_pairs'], len(groups), dt, stats['pairs'] / max(dt, 1e-9)))
    return stats


def main_bulk(argv=None):
    p = argparse.ArgumentParser(description='Anchor Gene Fusion Detection (c) -- B200 anchoring stage')
    _common_flags(p)
    p.add_argument('--fastq1', type=str, default='fastq_1.fastq', help='The fastq1 file to scan')
    p.add_argument('--fastq2', type=str, default='fastq_2.fastq', help='The fastq2 file to scan')
    _tail_flags(p)
    args = p.parse_args(argv)
    gene_names = parse_gene_names(args.file_anchored_cds, args.gene_names)
    _mkdir(args.out_folder)
    # Under torchrun (one process per GPU) the read batches of the ONE FASTQ pair are dealt to the ranks,
    # every rank anchors its share on its own GPU and rank 0 writes the one set of files.
    rank = int(os.environ.get('RANK', '0'))

    def work_prefix(gene):      # <out>/<gene>_fusion/work_dir/<gene>_fusion   (Anchored_Fusion.py:127-131)
        folder = args.out_folder + '/' + gene + '_fusion'
        _mkdir(folder + '/work_dir/')
        _mkdir(folder + '/model_dir/')
        return folder + '/work_dir/' + gene + '_fusion'

    if rank == 0:
        fastas = split_anchor_fasta(args.file_anchored_cds, gene_names, lambda g: work_prefix(g) + '_anchored_gene_sequence.fa')
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        from .stage import _host_group
        import torch.distributed as dist
        _, _, group = _host_group()
        dist.barrier(group=group)           # the per-gene FASTA files exist before any rank reads them
    fastas = [work_prefix(g) + '_anchored_gene_sequence.fa' for g in gene_names]
    todo = [(g, fa) for g, fa in zip(gene_names, fastas)
            if not (os.path.exists(work_prefix(g) + '_anchored_reads.bam') and os.path.exists(work_prefix(g) + '_realign_reads.bam'))]
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch.distributed as dist
        dist.barrier(group=group)           # every rank decided on the same list before rank 0 starts writing
    for g in gene_names:
        if g not in [t[0] for t in todo] and rank == 0:
            print('[anchoring] %s: outputs exist, skipping (same existence guard as the reference)' % work_prefix(g))
    if len(todo) == 1:
        run_gene_sample(todo[0][1], todo[0][0], args.fastq1, args.fastq2, work_prefix(todo[0][0]), args)
    elif todo:
        # several anchored genes: decode and pack the FASTQ pair once, copy every batch to the GPU once and
        # scan it for every gene while it is resident (the reference re-reads both files once per gene,
        # Anchored_Fusion.py:126,182)
        t0 = time.time()
        gas = [GeneAnchorer(fa, args.gpu_number, g) for g, fa in todo]
        all_stats = anchor_stage_multi(gas, args.fastq1, args.fastq2, [work_prefix(g) for g, _ in todo], thread=args.thread)
        if all_stats is not None:
            for (g, _), stats in zip(todo, all_stats):
                groups = write_split_points(stats, g, work_prefix(g) + '_split_points.txt')
                contiguity_stage(work_prefix(g), g, args)
                print('[anchoring] %s: %d pairs, %d anchored reads, %d half-anchored pairs, %d split-point groups'
                      % (g, stats['pairs'], stats['anchored'], stats['half_anchored_pairs'], len(groups)))
            print('[anchoring] %d genes in one pass over the reads, %.2f s' % (len(todo), time.time() - t0))
    return 0


def discover_cells(fastq_dir):
    """(cell, file_1, file_2) triples as the reference finds them (Anchored_Fusion_singlecell.py:86-113):
    sorted directory listing, `<cell>_1.<ext>` immediately followed by `<cell>_2.<ext>`."""
    names = sorted(os.listdir(fastq_dir + '/'))
    cells = []
    for i, f in enumerate(names):
        for ext in ('.fastq', '.fastq.gz', '.fq.gz', '.fq'):
            m = re.findall(r'(\S+)_1' + re.escape(ext) + '$', f)
            if m:
                nxt = names[i + 1] if i + 1 < len(names) else ''
                if nxt == m[0] + '_2' + ext:
                    cells.append((m[0], f, nxt))
                break
    return cells


def main_singlecell(argv=None):
    p = argparse.ArgumentParser(description='Anchor Gene Fusion Detection, single cell (c) -- B200 anchoring stage')
    _common_flags(p)
    p.add_argument('--fastq_dir', type=str, required=True, default='', help='The folder of fastq files, one <cell>_1/<cell>_2 pair per cell')
    _tail_flags(p)
    args = p.parse_args(argv)
    gene_names = parse_gene_names(args.file_anchored_cds, args.gene_names)
    cells = discover_cells(args.fastq_dir)
    _mkdir(args.out_folder)

    def gene_prefix(gene):      # <out>/<gene>/work_dir/<gene>_fusion   (Anchored_Fusion_singlecell.py:156-160)
        folder = args.out_folder + '/' + gene
        _mkdir(folder + '/work_dir/')
        _mkdir(folder + '/model_dir/')
        return folder + '/work_dir/' + gene + '_fusion'

    fastas = split_anchor_fasta(args.file_anchored_cds, gene_names, lambda g: gene_prefix(g) + '_anchored_gene_sequence.fa')
    # Cells are independent and every output is a per-cell file, so a multi-GPU run is one process per
    # GPU (torchrun) with the cells dealt round-robin to the ranks: no exchange at all.
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    gas = [GeneAnchorer(fa, args.gpu_number, gene) for gene, fa in zip(gene_names, fastas)]   # one index upload + staging per gene, not per cell

    def prefix_of(gene, cell):  # <out>/<gene>/work_dir/<cell>/<gene>_fusion   (Anchored_Fusion_singlecell.py:208)
        return args.out_folder + '/' + gene + '/work_dir/' + cell + '/' + gene + '_fusion'

    t0 = time.time()
    mine = []
    for ci, (cell, f1, f2) in enumerate(cells):
        if ci % world != rank:
            continue
        for gene in gene_names:
            _mkdir(args.out_folder + '/' + gene + '/work_dir/' + cell)
        # a cell is skipped only when every gene's outputs exist (same existence guards as the reference)
        if all(os.path.exists(prefix_of(g, cell) + '_anchored_reads.bam') and os.path.exists(prefix_of(g, cell) + '_realign_reads.bam')
               for g in gene_names):
            continue
        mine.append((cell, args.fastq_dir + '/' + f1, args.fastq_dir + '/' + f2))
    n_pairs = n_cells = 0
    if mine:
        # All of this rank's cells go through ONE reader: their files are decoded concurrently, packed back to
        # back into shared batches (1 M pairs per GPU pass instead of one pass per 5 k-pair cell) and every
        # batch is scanned for all genes while resident; the hit lists are split back per cell on the host.
        batcher = ContiguityBatcher(args)          # with --file_ref_seq: the genome pass, shared by many cells per call

        def on_cell(cell, cell_stats):
            for gene, st in zip(gene_names, cell_stats):
                write_split_points(st, gene, prefix_of(gene, cell) + '_split_points.txt')
                batcher.add(prefix_of(gene, cell), gene)

        res = anchor_cells(gas, mine, prefix_of, thread=args.thread, on_cell=on_cell)
        batcher.flush()
        n_pairs, n_cells = res['pairs'], res['cells']
        print('[anchoring] rank %d/%d: scan %.2f s, per-cell files %.2f s (%.2f ms per cell), %d reader threads'
              % (rank, world, res['seconds_scan'], res['seconds_write'], 1e3 * res['seconds_write'] / max(n_cells, 1), res['threads']))
    dt = time.time() - t0
    print('[anchoring] rank %d/%d: %d cells, %d pairs, %d genes, %.2f s (%.0f pairs/s)'
          % (rank, world, n_cells, n_pairs, len(gene_names), dt, n_pairs / max(dt, 1e-9)))
    return 0


if __name__ == '__main__':
    sys.exit(main_bulk())
