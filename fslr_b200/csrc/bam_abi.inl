// C ABI of the BAM -> mappings table producer (kernels: bam.cuh).  Included by fslr_b200.cu inside its extern "C" block.
}  // extern "C"
struct BamState {
    std::vector<void *> allocs;
    unsigned char *text; long long n;
    int M, NR, N, n_ref, n_primers, with_regions, ov_float;
    bam::Recs R;
    bam::Rows D;
    int *short_anchor, *name_rec, *first_rec, *rank;   // rank[rid in first-appearance order] = read id in output order
};
template <typename T>
static int bam_palloc(fslrc_ctx *ctx, T **p, int64_t n) {               // persistent (until fslrc_bam_close)
    void *q = nullptr;
    CK(cudaMallocAsync(&q, (size_t)(n > 0 ? n : 1) * sizeof(T), ctx->stream));
    ctx->bam->allocs.push_back(q);
    *p = (T *)q;
    return 0;
}
#define BPA(ptr, n) do { int r__ = bam_palloc(ctx, &(ptr), (int64_t)(n)); if (r__) return r__; } while (0)
static void bam_free(fslrc_ctx *ctx) {
    if (!ctx->bam) return;
    for (void *p : ctx->bam->allocs) cudaFreeAsync(p, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    delete ctx->bam; ctx->bam = nullptr;
}
static int bam_alloc_rows(fslrc_ctx *ctx, bam::Rows *W, int n, bool persistent) {
    int **f[15] = {&W->rid, &W->chrom, &W->rstart, &W->rend, &W->naln, &W->aln, &W->qstart, &W->qend, &W->strand, &W->mapq, &W->qlen,
                   &W->as, &W->inferred, &W->overlaps, &W->seq_rec};
    for (int k = 0; k < 15; k++) {
        if (persistent) BPA(*f[k], n); else DA(*f[k], n);
    }
    return 0;
}
static int bam_fail(fslrc_ctx *ctx, int code, const char *msg) {
    free_all(ctx); bam_free(ctx);
    return fail(ctx, code, "%s", msg);
}
extern "C" {

static int bam_primers(fslrc_ctx *ctx, bam::Primers &pr, const char *primer_names, const int32_t *primer_seq_len, int32_t n_primers) {
    memset(&pr, 0, sizeof(pr));
    pr.n = n_primers;
    const char *p = primer_names; int o = 0;
    for (int k = 0; k < n_primers; k++) {
        const int l = (int)strlen(p);
        if (o + l > (int)sizeof(pr.names)) return fail(ctx, FSLRC_ERR_ARG, "bam: primer names too long");
        memcpy(pr.names + o, p, l);
        pr.off[k] = o; pr.len[k] = primer_seq_len[k];
        o += l; p += l + 1;
    }
    pr.off[n_primers] = o;
    return 0;
}
static int bam_check_args(fslrc_ctx *ctx, const void *text, const void *info, int64_t n_bytes, int64_t first_record, int32_t n_ref,
                          const char *primer_names, const int32_t *primer_seq_len, int32_t n_primers, const int32_t *region_chrom,
                          const int32_t *region_start, const int32_t *region_end, int32_t n_regions) {
    if (!text || !info || n_bytes <= 0 || first_record < 0 || n_ref < 0) return fail(ctx, FSLRC_ERR_ARG, "bam: null or empty input");
    if (n_primers < 0 || n_primers > bam::MAX_PRIMERS || (n_primers > 0 && (!primer_names || !primer_seq_len)))
        return fail(ctx, FSLRC_ERR_ARG, "bam: at most 64 primers");
    if (n_regions > 0 && (!region_chrom || !region_start || !region_end)) return fail(ctx, FSLRC_ERR_ARG, "bam: null regions");
    return 0;
}
// record boundaries on the host: a chain of block_size fields; unmapped records (flag 4, :25) are dropped here
static int bam_host_walk(fslrc_ctx *ctx, const uint8_t *text, int64_t n_bytes, int64_t first_record, std::vector<long long> &off, int64_t *n_records) {
    int64_t p = first_record;
    *n_records = 0;
    while (p + 4 <= n_bytes) {
        int32_t bs; memcpy(&bs, text + p, 4);
        if (bs < 32 || p + 4 + (int64_t)bs > n_bytes) return fail(ctx, FSLRC_ERR_ARG, "bam: truncated or corrupt alignment record");
        uint16_t flag; memcpy(&flag, text + p + 18, 2);
        if (!(flag & 4)) off.push_back(p);
        (*n_records)++;
        p += 4 + (int64_t)bs;
    }
    if (p != n_bytes) return fail(ctx, FSLRC_ERR_ARG, "bam: trailing bytes after the last alignment record");
    if (off.size() > 0x7ffffff0ull / 2) return fail(ctx, FSLRC_ERR_ARG, "bam: too many records");
    return 0;
}
// everything after the record offsets are on the device: parse, group, primary, rows, sorts (see bam.cuh)
static int bam_build(fslrc_ctx *ctx, BamState *B, Pipe *P, const long long *rec_off, int M, int64_t n_bytes, int32_t n_ref,
                     const bam::Primers &pr, int32_t n_primers, const int32_t *region_chrom, const int32_t *region_start,
                     const int32_t *region_end, int32_t n_regions, uint64_t hash_seed, fslrc_bam_info *info) {
    cudaStream_t st = ctx->stream;
    B->M = M; B->n = n_bytes; B->n_ref = n_ref; B->n_primers = n_primers; B->with_regions = n_regions >= 0; B->ov_float = 0;
    const int TB = 256;
    bam::Recs &R = B->R;
    BPA(R.flag, M); BPA(R.ref, M); BPA(R.pos1, M); BPA(R.rend, M); BPA(R.mapq, M); BPA(R.qs, M); BPA(R.qe, M); BPA(R.qlen, M);
    BPA(R.as, M); BPA(R.lseq, M); BPA(R.nlen, M); BPA(R.seq_off, M); BPA(R.noff, M); BPA(R.hash, M);
    int *err, *dmax; int64_t *dcount;
    DA(err, 1); DA(dmax, 2); DA(dcount, 4);
    CK(cudaMemsetAsync(err, 0, sizeof(int), st)); CK(cudaMemsetAsync(dmax, 0, 2 * sizeof(int), st));
    CK(cudaMemsetAsync(dcount, 0, 4 * sizeof(int64_t), st));
    KL(bam::k_bam_parse, nblk(M, TB), TB, B->text, (long long)n_bytes, rec_off, M, (int)n_ref, (unsigned long long)hash_seed, R, dmax, err);
    // ---- read ids in order of first appearance (the dict of :23-26), records grouped by read in file order
    int *rid; DA(rid, M);
    BPA(B->first_rec, M);
    { int r = tsv_intern(ctx, P, B->text, M, R.noff, R.nlen, R.hash, rid, B->first_rec, err, dcount); if (r) return r; }
    CK(cudaMemcpyAsync(ctx->h_pin, dcount, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_pin + 8, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_pin + 9, dmax, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int e = (int)(ctx->h_pin[8] & 0xffffffff);
    const int maxlen = (int)(ctx->h_pin[9] & 0xffffffff);
    const int NR = B->NR = (int)ctx->h_pin[0];
#define BAM_ERRS(e)                                                                                                              \
    do {                                                                                                                         \
        if ((e) & bam::BE_COLLISION) { free_all(ctx); bam_free(ctx); return fail(ctx, FSLRC_ERR_HASH_COLLISION, "bam: two read names share a 64-bit hash; retry with another hash_seed"); } \
        if ((e) & bam::BE_TRUNC) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: alignment record shorter than its fields");           \
        if ((e) & bam::BE_NOCIGAR) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: mapped record without CIGAR (collect_mapping_info.py:12)"); \
        if ((e) & bam::BE_AUX) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: malformed aux block or non-integer AS tag");           \
        if ((e) & bam::BE_NOAS) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: record without AS tag (collect_mapping_info.py:44,88)"); \
        if ((e) & bam::BE_RANGE) return bam_fail(ctx, FSLRC_ERR_RANGE, "bam: coordinate or score beyond int32, or a mapped record whose reference id is not in the header");                 \
        if ((e) & bam::BE_NOPRIMARY) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: read without a primary record (collect_mapping_info.py:46-48)"); \
        if ((e) & bam::BE_NOSEQ) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: primary record without sequence (collect_mapping_info.py:101-103)"); \
        if ((e) & bam::BE_NAME) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: single-alignment read whose name does not end in <primer>_<primer> (collect_mapping_info.py:112-113)"); \
        if ((e) & bam::BE_PRIMER) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: primer named in a read name is not in the primer table (collect_mapping_info.py:133,151)"); \
    } while (0)
    BAM_ERRS(e);
    int *iotaM, *g, *rd_start, *rd_end; unsigned *ks;
    DA(iotaM, M); DA(g, M); DA(ks, M); DA(rd_start, NR); DA(rd_end, NR);
    KL(k_iota, nblk(M, TB), TB, iotaM, M);
    { int r = sort_pairs(ctx, P, (const unsigned *)rid, ks, iotaM, g, M, 0, bits_for(NR)); if (r) return r; }
    KL(bam::k_bam_bounds, nblk(M, TB), TB, M, ks, rd_start, rd_end);
    int *rd_naln, *rd_nrows, *rd_pri, *rd_row0; int4 *rd_inf;
    DA(rd_naln, NR); DA(rd_nrows, NR); DA(rd_pri, NR); DA(rd_row0, NR); DA(rd_inf, NR);
    KL(bam::k_bam_reads, nblk(NR, TB), TB, NR, B->text, g, rd_start, rd_end, R, pr, rd_naln, rd_nrows, rd_pri, rd_inf, dmax + 1, err);
    { int r = xscan(ctx, P, rd_nrows, rd_row0, NR, dcount + 1); if (r) return r; }
    CK(cudaMemcpyAsync(ctx->h_pin, dcount, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_pin + 8, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_pin + 9, dmax, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    e = (int)(ctx->h_pin[8] & 0xffffffff);
    BAM_ERRS(e);
#undef BAM_ERRS
    const int maxnaln = (int)(ctx->h_pin[9] >> 32);
    if (ctx->h_pin[1] > 0x7ffffff0LL / 2) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: too many rows");
    const int N = B->N = (int)ctx->h_pin[1];
    B->ov_float = B->with_regions && N > M;
    // ---- rows in `res` order
    bam::Rows W;
    { int r = bam_alloc_rows(ctx, &W, N, false); if (r) return r; }
    bam::Regions G; G.n = n_regions > 0 ? n_regions : 0; G.chrom = G.start = G.end = nullptr;
    if (G.n > 0) {
        int *gc, *gs, *ge; DA(gc, G.n); DA(gs, G.n); DA(ge, G.n);
        CK(cudaMemcpyAsync(gc, region_chrom, sizeof(int) * G.n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(gs, region_start, sizeof(int) * G.n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ge, region_end, sizeof(int) * G.n, cudaMemcpyHostToDevice, st));
        G.chrom = gc; G.start = gs; G.end = ge;
    }
    KL(bam::k_bam_rows, nblk(NR, TB), TB, NR, g, rd_start, rd_end, rd_row0, rd_naln, rd_pri, rd_inf, R, G, n_ref, W);
    // ---- reads by (n_alignments desc, qname asc): LSD string sort, 4 name bytes per pass, then one pass on n_alignments
    int *pa, *pb, *iotaR; unsigned *ka, *kb;
    DA(pa, NR); DA(pb, NR); DA(iotaR, NR); DA(ka, std::max(NR, N)); DA(kb, std::max(NR, N));
    KL(k_iota, nblk(NR, TB), TB, iotaR, NR);
    const int *cur = iotaR;
    for (int c = (maxlen + 3) / 4 - 1; c >= 0; c--) {
        int *dst = cur == pa ? pb : pa;
        KL(bam::k_bam_namekey, nblk(NR, TB), TB, NR, B->text, cur, B->first_rec, R.noff, R.nlen, c, ka);
        int r = sort_pairs(ctx, P, ka, kb, cur, dst, NR, 0, 32); if (r) return r;
        cur = dst;
    }
    {
        int *dst = cur == pa ? pb : pa;
        KL(bam::k_bam_gather_key, nblk(NR, TB), TB, NR, cur, rd_naln, maxnaln, -1, 0u, ka);
        int r = sort_pairs(ctx, P, ka, kb, cur, dst, NR, 0, bits_for((int64_t)maxnaln + 1)); if (r) return r;
        cur = dst;
    }
    BPA(B->rank, NR);
    KL(bam::k_bam_invert, nblk(NR, TB), TB, NR, cur, B->rank);
    // ---- rows by (read rank, qstart, res order): two stable sorts from res order (:163,174)
    int *iotaN, *v1, *order;
    DA(iotaN, N); DA(v1, N); DA(order, N);
    KL(k_iota, nblk(N, TB), TB, iotaN, N);
    KL(bam::k_bam_gather_key, nblk(N, TB), TB, N, iotaN, W.qstart, 0, 1, 0x80000000u, ka);
    { int r = sort_pairs(ctx, P, ka, kb, iotaN, v1, N, 0, 32); if (r) return r; }
    KL(bam::k_bam_rowkey, nblk(N, TB), TB, N, v1, W.rid, B->rank, ka);
    { int r = sort_pairs(ctx, P, ka, kb, v1, order, N, 0, bits_for(NR)); if (r) return r; }
    int *fa, *la; DA(fa, NR); DA(la, NR);
    KL(bam::k_bam_anchor, nblk(N, TB), TB, N, order, W.rid, W.aln, fa, la);
    { int r = bam_alloc_rows(ctx, &B->D, N, true); if (r) return r; }
    BPA(B->short_anchor, N); BPA(B->name_rec, N);
    KL(bam::k_bam_final, nblk(N, TB), TB, N, order, W, fa, la, B->rank, B->first_rec, B->D, B->short_anchor, B->name_rec);
    CK(cudaEventRecord(ctx->ev[2], st));
    CK(cudaStreamSynchronize(st));
    { cudaError_t ce = cudaGetLastError(); if (ce != cudaSuccess) { free_all(ctx); bam_free(ctx); return fail(ctx, FSLRC_ERR_CUDA, "bam: %s", cudaGetErrorString(ce)); } }
    free_all(ctx);
    float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
    info->n_mapped = M; info->n_reads = NR; info->n_rows = N; info->n_chrom = n_ref + n_primers; info->overlaps_as_float = B->ov_float;
    info->read_id = B->D.rid; info->chrom = B->D.chrom; info->rstart = B->D.rstart; info->rend = B->D.rend; info->n_alignments = B->D.naln;
    info->aln_size = B->D.aln; info->qstart = B->D.qstart; info->qend = B->D.qend; info->strand = B->D.strand; info->mapq = B->D.mapq;
    info->qlen = B->D.qlen; info->alignment_score = B->D.as; info->short_anchor = B->short_anchor; info->inferred_by_primer = B->D.inferred;
    info->overlaps_region = B->D.overlaps; info->parse_ms = ms;
    return 0;
}

int fslrc_bam_open(fslrc_ctx *ctx, const uint8_t *text, int64_t n_bytes, int64_t first_record, int32_t n_ref,
                   const char *primer_names, const int32_t *primer_seq_len, int32_t n_primers,
                   const int32_t *region_chrom, const int32_t *region_start, const int32_t *region_end, int32_t n_regions,
                   uint64_t hash_seed, fslrc_bam_info *info, void *stream) {
    if (!ctx) return FSLRC_ERR_ARG;
    { int r = bam_check_args(ctx, text, info, n_bytes, first_record, n_ref, primer_names, primer_seq_len, n_primers, region_chrom,
                             region_start, region_end, n_regions); if (r) return r; }
    if (first_record > n_bytes) return fail(ctx, FSLRC_ERR_ARG, "bam: null or empty input");
    CK(cudaSetDevice(ctx->device));
    ctx->stream = (cudaStream_t)stream;
    cudaStream_t st = ctx->stream;
    bam_free(ctx);
    memset(info, 0, sizeof(*info));
    bam::Primers pr;
    { int r = bam_primers(ctx, pr, primer_names, primer_seq_len, n_primers); if (r) return r; }
    std::vector<long long> off;
    int64_t n_records = 0;
    { int r = bam_host_walk(ctx, text, n_bytes, first_record, off, &n_records); if (r) return r; }
    const int M = (int)off.size();
    info->n_records = n_records; info->n_mapped = M;
    if (M == 0) return fail(ctx, FSLRC_ERR_ARG, "bam: no mapped records (the reference fails on an empty table, collect_mapping_info.py:163)");
    BamState *B = ctx->bam = new BamState();
    memset((void *)&B->R, 0, sizeof(B->R)); memset((void *)&B->D, 0, sizeof(B->D));
    Pipe Pp; memset((void *)&Pp, 0, sizeof(Pp)); Pipe *P = &Pp;       // (scratch of the scan / sort primitives)
    B->n = n_bytes;
    BPA(B->text, n_bytes + 64);
    CK(cudaMemcpyAsync(B->text, text, n_bytes, cudaMemcpyHostToDevice, st));
    long long *rec_off; DA(rec_off, M);
    CK(cudaMemcpyAsync(rec_off, off.data(), sizeof(long long) * M, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[1], st));
    return bam_build(ctx, B, P, rec_off, M, n_bytes, n_ref, pr, n_primers, region_chrom, region_start, region_end, n_regions, hash_seed, info);
}

// The same from the COMPRESSED file: BGZF blocks are inflated on the device (inflate.cuh, one warp per block) and the record
// boundaries are found there too (bam_chain.cuh), so only the compressed bytes cross PCIe and the host touches nothing but
// the BGZF block headers.  `first_record` / `n_ref` come from the caller's parse of the BAM header (the first blocks).
int fslrc_bam_open_bgzf(fslrc_ctx *ctx, const uint8_t *file, int64_t n_bytes, int64_t first_record, int32_t n_ref,
                        const char *primer_names, const int32_t *primer_seq_len, int32_t n_primers,
                        const int32_t *region_chrom, const int32_t *region_start, const int32_t *region_end, int32_t n_regions,
                        uint64_t hash_seed, fslrc_bam_info *info, void *stream) {
    if (!ctx) return FSLRC_ERR_ARG;
    { int r = bam_check_args(ctx, file, info, n_bytes, first_record, n_ref, primer_names, primer_seq_len, n_primers, region_chrom,
                             region_start, region_end, n_regions); if (r) return r; }
    CK(cudaSetDevice(ctx->device));
    ctx->stream = (cudaStream_t)stream;
    cudaStream_t st = ctx->stream;
    bam_free(ctx);
    memset(info, 0, sizeof(*info));
    bam::Primers pr;
    { int r = bam_primers(ctx, pr, primer_names, primer_seq_len, n_primers); if (r) return r; }
    // ---- BGZF block table (host): gzip member header with the `BC` extra subfield (BSIZE), ISIZE in the trailer
    std::vector<long long> in_off, out_off; std::vector<int> in_len, out_len;
    int64_t total = 0;
    for (int64_t p = 0; p < n_bytes;) {
        if (n_bytes - p < 18 || file[p] != 0x1f || file[p + 1] != 0x8b || file[p + 2] != 8 || !(file[p + 3] & 4))
            return fail(ctx, FSLRC_ERR_ARG, "bam: not a BGZF file");
        const int xlen = file[p + 10] | (file[p + 11] << 8);
        int64_t q = p + 12, bsize = -1;
        while (q + 4 <= p + 12 + xlen && q + 4 <= n_bytes) {
            const int slen = file[q + 2] | (file[q + 3] << 8);
            if (file[q] == 66 && file[q + 1] == 67 && slen == 2 && q + 6 <= n_bytes) bsize = (file[q + 4] | (file[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        if (bsize < 12 + xlen + 8 || p + bsize > n_bytes) return fail(ctx, FSLRC_ERR_ARG, "bam: truncated BGZF block");
        uint32_t isize; memcpy(&isize, file + p + bsize - 4, 4);
        if (isize > 65536u) return fail(ctx, FSLRC_ERR_ARG, "bam: BGZF block larger than 64 KiB");
        in_off.push_back(p + 12 + xlen); in_len.push_back((int)(bsize - xlen - 20)); out_off.push_back(total); out_len.push_back((int)isize);
        total += isize;
        p += bsize;
    }
    if (first_record > total) return fail(ctx, FSLRC_ERR_ARG, "bam: first_record beyond the stream");
    if (in_off.size() > 0x7ffffff0ull / 64) return fail(ctx, FSLRC_ERR_ARG, "bam: too many BGZF blocks");
    const int nb = (int)in_off.size();
    const int64_t n_tiles64 = (total + bam::TILE - 1) / bam::TILE;
    const int n_tiles = (int)n_tiles64;
    if (total <= 0 || n_tiles <= 0) return fail(ctx, FSLRC_ERR_ARG, "bam: empty stream");
    BamState *B = ctx->bam = new BamState();
    memset((void *)&B->R, 0, sizeof(B->R)); memset((void *)&B->D, 0, sizeof(B->D));
    Pipe Pp; memset((void *)&Pp, 0, sizeof(Pp)); Pipe *P = &Pp;
    B->n = total;
    BPA(B->text, total + 64);
    unsigned char *d_file; long long *d_in_off, *d_out_off; int *d_in_len, *d_out_len, *ierr;
    DA(d_file, n_bytes); DA(d_in_off, nb); DA(d_out_off, nb); DA(d_in_len, nb); DA(d_out_len, nb); DA(ierr, 2);
    CK(cudaMemcpyAsync(d_file, file, n_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_in_off, in_off.data(), sizeof(long long) * nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_out_off, out_off.data(), sizeof(long long) * nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_in_len, in_len.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_out_len, out_len.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ierr, 0, 2 * sizeof(int), st));
    CK(cudaEventRecord(ctx->ev[1], st));
    KL(inflate::k_inflate, nblk(nb, inflate::INF_WARPS), inflate::INF_WARPS * 32, d_file, d_in_off, d_in_len, d_out_off, d_out_len, nb, B->text, ierr);
    // ---- record boundaries: tiles of the inflated stream
    const int TB = 256;
    long long *tfirst, *texit; int *tcount, *tbase; int64_t *tot; unsigned long long *nrec;
    DA(tfirst, n_tiles); DA(texit, n_tiles); DA(tcount, n_tiles); DA(tbase, n_tiles); DA(tot, 2);
    nrec = (unsigned long long *)(tot + 1);
    CK(cudaMemsetAsync(tot, 0, 2 * sizeof(int64_t), st));
    KL(bam::k_bam_tile_first, nblk((int64_t)n_tiles * 32, TB), TB, B->text, (long long)total, (long long)first_record, n_ref, n_tiles, tfirst);
    KL(bam::k_bam_tile_walk, nblk(n_tiles, 64), 64, B->text, (long long)total, n_tiles, tfirst, tcount, nrec, texit, (const int *)nullptr, (long long *)nullptr);
    KL(bam::k_bam_tile_check, nblk(n_tiles, TB), TB, n_tiles, (long long)total, (long long)first_record, tfirst, texit, ierr + 1);
    { int r = xscan(ctx, P, tcount, tbase, n_tiles, tot); if (r) return r; }
    CK(cudaMemcpyAsync(ctx->h_pin, tot, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_pin + 8, ierr, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int inf_err = (int)(ctx->h_pin[8] & 0xffffffff);
    const int chain_bad = (int)(ctx->h_pin[8] >> 32) || getenv("FSLRC_BAM_FORCE_HOST_WALK") != nullptr;   // (the variable: test hook for the fallback)
    if (inf_err) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: corrupt DEFLATE data in a BGZF block");
    long long *rec_off = nullptr;
    int M = 0;
    if (!chain_bad) {
        if (ctx->h_pin[0] > 0x7ffffff0LL / 2) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: too many records");
        M = (int)ctx->h_pin[0];
        info->n_records = (int64_t)ctx->h_pin[1]; info->n_mapped = M;
        if (M == 0) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: no mapped records (the reference fails on an empty table, collect_mapping_info.py:163)");
        DA(rec_off, M);
        KL(bam::k_bam_tile_walk, nblk(n_tiles, 64), 64, B->text, (long long)total, n_tiles, tfirst, tcount, nrec, texit, tbase, rec_off);
    } else {                                                   // a tile's guess was off the chain: walk on the host instead
        std::vector<uint8_t> host((size_t)total);
        CK(cudaMemcpyAsync(host.data(), B->text, total, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        std::vector<long long> off; int64_t n_records = 0;
        { int r = bam_host_walk(ctx, host.data(), total, first_record, off, &n_records); if (r) { free_all(ctx); bam_free(ctx); return r; } }
        M = (int)off.size();
        info->n_records = n_records; info->n_mapped = M;
        if (M == 0) return bam_fail(ctx, FSLRC_ERR_ARG, "bam: no mapped records (the reference fails on an empty table, collect_mapping_info.py:163)");
        DA(rec_off, M);
        CK(cudaMemcpyAsync(rec_off, off.data(), sizeof(long long) * M, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    info->reserved = chain_bad;                                // 1: the host fallback ran
    return bam_build(ctx, B, P, rec_off, M, total, n_ref, pr, n_primers, region_chrom, region_start, region_end, n_regions, hash_seed, info);
}

// the inflated BAM stream back on the host (read names live there): *n receives its length; out == NULL only asks for it
int fslrc_bam_read_stream(fslrc_ctx *ctx, uint8_t *out, int64_t cap, int64_t *n) {
    if (!ctx || !ctx->bam || !n) return FSLRC_ERR_ARG;
    BamState *B = ctx->bam;
    CK(cudaSetDevice(ctx->device));
    *n = B->n;
    if (!out) return 0;
    if (cap < B->n) return fail(ctx, FSLRC_ERR_ARG, "bam: output buffer too small");
    CK(cudaMemcpyAsync(out, B->text, B->n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fslrc_bam_write_mappings_bed(fslrc_ctx *ctx, const char *chrom_names, const char *fslr_version, char *out, int64_t cap,
                                 int64_t *n_out, void *stream) {
    if (!ctx || !ctx->bam || !chrom_names || !fslr_version || !n_out) return FSLRC_ERR_ARG;
    BamState *B = ctx->bam;
    CK(cudaSetDevice(ctx->device));
    ctx->stream = (cudaStream_t)stream;
    cudaStream_t st = ctx->stream;
    Pipe Pp; memset((void *)&Pp, 0, sizeof(Pp)); Pipe *P = &Pp;
    const int N = B->N, TB = 256, NC = B->n_ref + B->n_primers;
    bam::Emit E; memset((void *)&E, 0, sizeof(E));
    E.ver_len = (int)strlen(fslr_version);
    if (E.ver_len > (int)sizeof(E.ver)) return fail(ctx, FSLRC_ERR_ARG, "bam: version string too long");
    memcpy(E.ver, fslr_version, E.ver_len);
    std::vector<int> coff(NC + 1); std::string cat;
    { const char *p = chrom_names; for (int c = 0; c < NC; c++) { coff[c] = (int)cat.size(); const size_t l = strlen(p); cat.append(p, l); p += l + 1; } coff[NC] = (int)cat.size(); }
    char *d_names; int *d_coff;
    DA(d_names, cat.size() + 1); DA(d_coff, NC + 1);
    CK(cudaMemcpyAsync(d_names, cat.data(), cat.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_coff, coff.data(), sizeof(int) * (NC + 1), cudaMemcpyHostToDevice, st));
    E.D = B->D; E.short_anchor = B->short_anchor; E.name_rec = B->name_rec; E.R = B->R; E.C.text = d_names; E.C.off = d_coff;
    E.with_regions = B->with_regions; E.ov_float = B->ov_float;
    std::string header = "chrom\trstart\trend\tqname\tn_alignments\taln_size\tqstart\tqend\tstrand\tmapq\tqlen\talignment_score\t"
                         "short_anchor<50bp\tfslr_version\tinferred_by_primer\tseq";
    if (B->with_regions) header += "\toverlaps_region";
    header += "\n";
    long long *len, *off; int64_t *tot;
    DA(len, N); DA(off, N); DA(tot, 1);
    KL(bam::k_bam_outlen, nblk(N, TB), TB, N, E, len);
    {
        const int tiles = nblk(N, prims::SC_TILE);
        int r = prim_scratch(ctx, P, sizeof(unsigned long long) * tiles); if (r) return r;
        KL(prims::k_scan_excl<long long>, tiles, prims::SC_THREADS, len, off, N, (unsigned long long *)(P->prim + 256), (unsigned *)P->prim, (long long *)tot);
    }
    CK(cudaMemcpyAsync(ctx->h_pin, tot, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_out = ctx->h_pin[0] + (int64_t)header.size();
    int rc = 0;
    if (out) {
        if (cap < *n_out) rc = fail(ctx, FSLRC_ERR_ARG, "bam: output buffer too small");
        else {
            unsigned char *d_out; DA(d_out, *n_out);
            CK(cudaMemcpyAsync(d_out, header.data(), header.size(), cudaMemcpyHostToDevice, st));
            KL(bam::k_bam_emit, nblk((int64_t)N * 32, TB), TB, N, E, B->text, off, (long long)header.size(), d_out);
            CK(cudaMemcpyAsync(out, d_out, *n_out, cudaMemcpyDeviceToHost, st));
        }
    }
    CK(cudaStreamSynchronize(st));
    free_all(ctx);
    CK(cudaStreamSynchronize(st));
    return rc;
}

int fslrc_bam_read_names(fslrc_ctx *ctx, int64_t *offsets, int32_t *lengths) {
    if (!ctx || !ctx->bam || !offsets || !lengths) return FSLRC_ERR_ARG;
    BamState *B = ctx->bam;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int NR = B->NR, M = B->M;
    std::vector<int> fr(NR), rk(NR), nl(M); std::vector<long long> no(M);
    CK(cudaMemcpyAsync(fr.data(), B->first_rec, sizeof(int) * NR, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(rk.data(), B->rank, sizeof(int) * NR, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nl.data(), B->R.nlen, sizeof(int) * M, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(no.data(), B->R.noff, sizeof(long long) * M, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int r = 0; r < NR; r++) { offsets[rk[r]] = no[fr[r]]; lengths[rk[r]] = nl[fr[r]]; }
    return 0;
}

void fslrc_bam_close(fslrc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    bam_free(ctx);
}
