// mg_host_demo -- drives GpuTrackExprScanner from a small text spec; used by tests/test_gpu_host_cpp.py to check the C++ host
// layer against the oracle, and as the example of a .Call-shaped caller. Not part of the library.
//
// spec (one directive per line):
//   chroms <n> <size0> ... <size n-1>
//   dense <name> <binsize> <file chrom0> ... (raw float32 bins per chromosome; "-" = no data for that chromosome)
//   sparse <name> <file chrom0> ...          (per chromosome: int64 n, then n starts, n ends (int64), n float32 values)
//   ivset <name> <file chrom0> ...           (per chromosome: int64 n, n starts, n ends)
//   vtrack <vname> dense|sparse|coverage <source name> <func id> <param> <sshift> <eshift>
//   seq <name> <file chrom0> ...             (ASCII bases per chromosome)
//   kmer <vname> <seq name> <mode 0 count / 1 frac> <kmer> <extend> <strand> <sshift> <eshift>
//   pwm <vname> <seq name> <mode> <pssm file: float64 L x 4> <prior> <bidirect> <extend> <strand> <score thresh> <sshift> <eshift>
//   scope <file>                             (int64 n, then n x {int64 chromid, start, end})
//   iterator <binsize> | iterator intervals <file>
//   buffer <rows>
//   exprs <vname> ...
//   out <file>      (int64 nrows, int64 nexprs, chrom int32[n], start int64[n], end int64[n], intervalID int32[n], nexprs x double[n])
//   summary <dense name> <binsize> <out file> (7 doubles)
//   group <ndev> <dense name> <binsize> <func id> <param> <sshift> <eshift> <out file>
//        the same track on a group of the first <ndev> devices (GpuGroupSession): gextract of the scope under the window, then gsummary,
//        gdist (breaks 0, 1, ..., 100, include.lowest) and gquantiles (0.1, 0.5, 0.9) of the plain track over the fixed-bin iterator
//        <binsize> -> int64 n, double[n] column, 7 doubles, 100 uint64, 3 doubles
#include <cstdio>
#include <cstring>
#include <tuple>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "GpuTrackExprScanner.h"

using namespace mishagpu;

template <typename T>
static std::vector<T> read_raw(const std::string &path, size_t offset_bytes = 0, int64_t count = -1)
{
	std::ifstream f(path, std::ios::binary);
	if (!f)
		throw std::runtime_error("cannot open " + path);
	f.seekg(0, std::ios::end);
	const size_t bytes = (size_t)f.tellg() - offset_bytes;
	f.seekg((std::streamoff)offset_bytes);
	std::vector<T> v(count < 0 ? bytes / sizeof(T) : (size_t)count);
	f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
	return v;
}

static std::vector<Interval> read_intervals(const std::string &path)
{
	const std::vector<int64_t> raw = read_raw<int64_t>(path);
	std::vector<Interval> out((size_t)raw[0]);
	for (size_t i = 0; i < out.size(); ++i)
		out[i] = Interval{(int32_t)raw[1 + 3 * i], raw[2 + 3 * i], raw[3 + 3 * i]};
	return out;
}

int main(int argc, char **argv)
{
	if (argc != 2) {
		std::fprintf(stderr, "usage: mg_host_demo <spec>\n");
		return 2;
	}
	try {
		std::ifstream spec(argv[1]);
		std::string line;
		mg_ctx *ctx = nullptr;
		std::vector<int64_t> sizes;
		std::map<std::string, mg_dense *> dense;
		std::map<std::string, std::pair<uint32_t, std::vector<std::string>>> dense_files;
		std::map<std::string, mg_sparse *> sparse;
		std::map<std::string, mg_ivset *> ivsets;
		std::vector<std::tuple<std::string, std::string, std::string, int, double, int64_t, int64_t>> vtracks;
		std::map<std::string, mg_seq *> seqs;
		struct PwmSpec {
			std::string name, seq;
			int mode;
			std::vector<double> pssm;
			double prior;
			int bidirect, extend, strand;
			float thresh;
			int64_t ss, es;
		};
		std::vector<PwmSpec> pwms;
		struct KmerSpec {
			std::string name, seq, kmer;
			int mode, extend, strand;
			int64_t ss, es;
		};
		std::vector<KmerSpec> kmers;
		std::vector<Interval> scope, iter_intervals;
		int64_t binsize = 0, buffer = int64_t(1) << 20;
		bool use_iter_intervals = false;
		std::vector<std::string> exprs;
		auto chk = [&](int rc) {
			if (rc)
				throw MishaGpuException(mg_last_error(ctx));
		};
		while (std::getline(spec, line)) {
			std::istringstream in(line);
			std::string cmd;
			if (!(in >> cmd) || cmd[0] == '#')
				continue;
			if (cmd == "chroms") {
				int n;
				in >> n;
				sizes.resize(n);
				for (auto &s : sizes)
					in >> s;
				if (mg_init(0, n, sizes.data(), &ctx))
					throw MishaGpuException(mg_last_error(nullptr));
			} else if (cmd == "dense") {
				std::string name, path;
				uint32_t bs;
				in >> name >> bs;
				mg_dense *t = nullptr;
				chk(mg_dense_create(ctx, bs, &t));
				dense_files[name].first = bs;
				for (size_t c = 0; c < sizes.size(); ++c) {
					in >> path;
					dense_files[name].second.push_back(path);
					if (path == "-")
						continue;
					const std::vector<float> bins = read_raw<float>(path);
					chk(mg_dense_stage(ctx, t, (int)c, bins.data(), (int64_t)bins.size()));
				}
				dense[name] = t;
			} else if (cmd == "sparse" || cmd == "ivset") {
				std::string name, path;
				in >> name;
				mg_sparse *t = nullptr;
				mg_ivset *s = nullptr;
				if (cmd == "sparse")
					chk(mg_sparse_create(ctx, &t));
				else
					chk(mg_ivset_create(ctx, &s));
				for (size_t c = 0; c < sizes.size(); ++c) {
					in >> path;
					const int64_t n = read_raw<int64_t>(path, 0, 1)[0];
					const std::vector<int64_t> st = read_raw<int64_t>(path, 8, n), en = read_raw<int64_t>(path, 8 + 8 * (size_t)n, n);
					if (t) {
						const std::vector<float> val = read_raw<float>(path, 8 + 16 * (size_t)n, n);
						chk(mg_sparse_stage(ctx, t, (int)c, st.data(), en.data(), val.data(), n));
					} else
						chk(mg_ivset_stage(ctx, s, (int)c, st.data(), en.data(), n));
				}
				if (t)
					sparse[name] = t;
				else
					ivsets[name] = s;
			} else if (cmd == "vtrack") {
				std::string v, kind, src;
				int func;
				double param;
				int64_t ss, es;
				in >> v >> kind >> src >> func >> param >> ss >> es;
				vtracks.emplace_back(v, kind, src, func, param, ss, es);
			} else if (cmd == "seq") {
				std::string name, path;
				in >> name;
				mg_seq *q = nullptr;
				chk(mg_seq_create(ctx, &q));
				for (size_t c = 0; c < sizes.size(); ++c) {
					in >> path;
					const std::vector<char> bases = read_raw<char>(path);
					chk(mg_seq_stage(ctx, q, (int)c, bases.data(), (int64_t)bases.size()));
				}
				seqs[name] = q;
			} else if (cmd == "pwm") {
				PwmSpec p;
				std::string path;
				in >> p.name >> p.seq >> p.mode >> path >> p.prior >> p.bidirect >> p.extend >> p.strand >> p.thresh >> p.ss >> p.es;
				p.pssm = read_raw<double>(path);
				pwms.push_back(p);
			} else if (cmd == "kmer") {
				KmerSpec k;
				in >> k.name >> k.seq >> k.mode >> k.kmer >> k.extend >> k.strand >> k.ss >> k.es;
				kmers.push_back(k);
			} else if (cmd == "scope") {
				std::string path;
				in >> path;
				scope = read_intervals(path);
			} else if (cmd == "iterator") {
				std::string a;
				in >> a;
				if (a == "intervals") {
					std::string path;
					in >> path;
					iter_intervals = read_intervals(path);
					use_iter_intervals = true;
				} else
					binsize = std::stoll(a);
			} else if (cmd == "buffer") {
				in >> buffer;
			} else if (cmd == "exprs") {
				std::string e;
				while (in >> e)
					exprs.push_back(e);
			} else if (cmd == "out") {
				std::string path;
				in >> path;
				GpuTrackExprScanner scanner(ctx, buffer);
				for (auto &vt : vtracks) {
					const std::string &kind = std::get<1>(vt), &src = std::get<2>(vt);
					if (kind == "dense")
						scanner.vtrack_dense(std::get<0>(vt), dense.at(src), std::get<3>(vt), std::get<4>(vt), std::get<5>(vt), std::get<6>(vt));
					else if (kind == "sparse")
						scanner.vtrack_sparse(std::get<0>(vt), sparse.at(src), std::get<3>(vt), std::get<4>(vt), std::get<5>(vt), std::get<6>(vt));
					else
						scanner.vtrack_coverage(std::get<0>(vt), ivsets.at(src), std::get<5>(vt), std::get<6>(vt));
				}
				for (auto &p : pwms)
					scanner.vtrack_pwm(p.name, seqs.at(p.seq), p.mode, p.pssm, p.prior, p.bidirect != 0, p.extend != 0, p.strand, p.thresh, p.ss, p.es);
				for (auto &k : kmers)
					scanner.vtrack_kmer(k.name, seqs.at(k.seq), k.mode, k.kmer, k.extend != 0, k.strand, k.ss, k.es);
				// both ways of consuming the scanner must agree: the reference-shaped per-row loop and whole columns
				const ExtractResult a = gextract(scanner, exprs, scope, binsize, use_iter_intervals ? &iter_intervals : nullptr, true);
				const ExtractResult b = gextract(scanner, exprs, scope, binsize, use_iter_intervals ? &iter_intervals : nullptr, false);
				if (a.chrom != b.chrom || a.start != b.start || a.end != b.end || a.interval_id != b.interval_id)
					throw std::runtime_error("per-row and whole-column scans disagree on the rows");
				for (size_t i = 0; i < exprs.size(); ++i)
					if (std::memcmp(a.values[i].data(), b.values[i].data(), a.values[i].size() * 8))
						throw std::runtime_error("per-row and whole-column scans disagree on " + exprs[i]);
				std::ofstream f(path, std::ios::binary);
				const int64_t n = (int64_t)a.chrom.size(), ne = (int64_t)exprs.size();
				f.write((const char *)&n, 8);
				f.write((const char *)&ne, 8);
				f.write((const char *)a.chrom.data(), n * 4);
				f.write((const char *)a.start.data(), n * 8);
				f.write((const char *)a.end.data(), n * 8);
				f.write((const char *)a.interval_id.data(), n * 4);
				for (auto &col : a.values)
					f.write((const char *)col.data(), n * 8);
			} else if (cmd == "summary") {
				std::string name, path;
				int64_t bs;
				in >> name >> bs >> path;
				double out[7];
				gsummary_dense(ctx, dense.at(name), scope, bs, MG_AVG, 0., 0, 0, out);
				std::ofstream f(path, std::ios::binary);
				f.write((const char *)out, sizeof(out));
			} else if (cmd == "group") {
				int ndev, func;
				std::string name, path;
				int64_t bs, ss, es;
				double param;
				in >> ndev >> name >> bs >> func >> param >> ss >> es >> path;
				std::vector<int> devs;
				for (int d = 0; d < ndev; ++d)
					devs.push_back(d);
				GpuGroupSession session(devs, sizes);
				const auto &files = dense_files.at(name);
				mg_gdense *gt = session.dense_create(files.first);
				for (size_t c = 0; c < files.second.size(); ++c)
					if (files.second[c] != "-") {
						const std::vector<float> bins = read_raw<float>(files.second[c]);
						session.dense_stage(gt, (int)c, bins.data(), (int64_t)bins.size());
					}
				const ExtractResult r = session.gextract(gt, {{func, param}}, ss, es, scope);
				double sum7[7];
				session.gsummary(gt, scope, bs, MG_AVG, 0., 0, 0, sum7);
				std::vector<double> breaks;
				for (int b = 0; b <= 100; ++b)
					breaks.push_back((double)b);
				const std::vector<uint64_t> hist = session.gdist(gt, scope, bs, MG_AVG, 0., 0, 0, breaks, true);
				const std::vector<double> q = session.gquantiles(gt, scope, bs, MG_AVG, 0., 0, 0, {0.1, 0.5, 0.9});
				std::ofstream f(path, std::ios::binary);
				const int64_t n = (int64_t)r.values[0].size();
				f.write((const char *)&n, 8);
				f.write((const char *)r.values[0].data(), n * 8);
				f.write((const char *)sum7, sizeof(sum7));
				f.write((const char *)hist.data(), (std::streamsize)(hist.size() * 8));
				f.write((const char *)q.data(), (std::streamsize)(q.size() * 8));
			} else
				throw std::runtime_error("unknown directive " + cmd);
		}
		for (auto &kv : dense)
			mg_dense_free(ctx, kv.second);
		for (auto &kv : sparse)
			mg_sparse_free(ctx, kv.second);
		for (auto &kv : ivsets)
			mg_ivset_free(ctx, kv.second);
		for (auto &kv : seqs)
			mg_seq_free(ctx, kv.second);
		mg_shutdown(ctx);
	} catch (const std::exception &e) {
		std::fprintf(stderr, "mg_host_demo: %s\n", e.what());
		return 1;
	}
	return 0;
}
