// Reference harness (TEST INFRASTRUCTURE, not product code).
//
// Drives the UNMODIFIED reference sources under /root/reference/src (compiled by oracle/Makefile into
// oracle/_ref/) and dumps everything the parity tests need: lattice tables, frequency/cutoff meshes,
// vertex state, vertex flow, and the correlation measurements the reference would have written to HDF5.
//
// The reference headers only DECLARE SpinParser and CommandLineOptions (src/SpinParser.hpp:60-152,
// src/CommandLineOptions.hpp:17-90); their bodies live in SpinParser.cpp / CommandLineOptions.cpp, which need
// boost::program_options and MPI. This file supplies replacement bodies, which also makes it a `friend` of
// FrgCommon (src/FrgCommon.hpp:19) and FrgCore (src/FrgCore.hpp:31) exactly like the reference's own unit
// tests do (test/test_SU2VertexTwoParticle.cpp:6-20). The run loop below follows src/SpinParser.cpp:126-186.
//
// This TU is included at the end of oracle/harness/unity.cpp, after all reference translation units.

#include "SpinParser.hpp"
#include "FrgCore.hpp"
#include "FrgCoreFactory.hpp"
#include "LatticeModelFactory.hpp"
#include "TaskFileParser.hpp"
#include "SU2/SU2FrgCore.hpp"
#include "SU2/SU2EffectiveAction.hpp"
#include "XYZ/XYZFrgCore.hpp"
#include "XYZ/XYZEffectiveAction.hpp"
#include "TRI/TRIFrgCore.hpp"
#include "TRI/TRIEffectiveAction.hpp"

namespace harness
{
	typedef float real; // becomes double in the FP64 build (unity.cpp: #define float double)

	// ---- PFD container: [magic "PFD1"] then records {u32 nameLen, name, u8 dtype, u32 ndim, u64 dims[], raw data}
	enum DType : unsigned char { F32 = 0, F64 = 1, I32 = 2, I64 = 3, U8 = 4 };
	struct Writer
	{
		FILE *f = nullptr;
		bool open(const std::string &path) { f = fopen(path.c_str(), "wb"); if (!f) return false; fwrite("PFD1", 1, 4, f); return true; }
		void close() { if (f) fclose(f); f = nullptr; }
		void raw(const std::string &name, unsigned char dtype, const std::vector<uint64_t> &dims, const void *data, size_t bytes)
		{
			if (!f) return;
			uint32_t nl = (uint32_t)name.size(); fwrite(&nl, 4, 1, f); fwrite(name.data(), 1, nl, f);
			fwrite(&dtype, 1, 1, f);
			uint32_t nd = (uint32_t)dims.size(); fwrite(&nd, 4, 1, f);
			for (auto d : dims) fwrite(&d, 8, 1, f);
			fwrite(data, 1, bytes, f);
		}
		void reals(const std::string &name, const real *p, const std::vector<uint64_t> &dims)
		{
			size_t n = 1; for (auto d : dims) n *= d;
			raw(name, sizeof(real) == 8 ? F64 : F32, dims, p, n * sizeof(real));
		}
		void ints(const std::string &name, const std::vector<int> &v, std::vector<uint64_t> dims = {})
		{
			if (dims.empty()) dims = { (uint64_t)v.size() };
			raw(name, I32, dims, v.data(), v.size() * sizeof(int));
		}
		void doubles(const std::string &name, const std::vector<double> &v, std::vector<uint64_t> dims = {})
		{
			if (dims.empty()) dims = { (uint64_t)v.size() };
			raw(name, F64, dims, v.data(), v.size() * sizeof(double));
		}
		void scalar(const std::string &name, double v) { raw(name, F64, {}, &v, 8); }
	};

	// ---- minimal reader for --load-state (same container)
	struct Record { unsigned char dtype; std::vector<uint64_t> dims; std::vector<unsigned char> data; };
	inline std::map<std::string, Record> readPfd(const std::string &path)
	{
		std::map<std::string, Record> out;
		FILE *f = fopen(path.c_str(), "rb");
		if (!f) throw Exception(Exception::Type::IOError, "cannot open " + path);
		char magic[4]; if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "PFD1", 4) != 0) throw Exception(Exception::Type::IOError, "bad magic in " + path);
		while (true)
		{
			uint32_t nl; if (fread(&nl, 4, 1, f) != 1) break;
			std::string name(nl, ' '); if (fread(&name[0], 1, nl, f) != nl) break;
			Record r; if (fread(&r.dtype, 1, 1, f) != 1) break;
			uint32_t nd; if (fread(&nd, 4, 1, f) != 1) break;
			r.dims.resize(nd); size_t n = 1;
			for (uint32_t i = 0; i < nd; ++i) { if (fread(&r.dims[i], 8, 1, f) != 1) break; n *= r.dims[i]; }
			static const size_t es[] = { 4, 8, 4, 8, 1 };
			r.data.resize(n * es[r.dtype]);
			if (n && fread(r.data.data(), 1, r.data.size(), f) != r.data.size()) break;
			out[name] = r;
		}
		fclose(f);
		return out;
	}
	inline void loadReals(const Record &r, real *dst, size_t n)
	{
		size_t have = 1; for (auto d : r.dims) have *= d;
		if (have != n) throw Exception(Exception::Type::ArgumentError, "state array has wrong size");
		if (r.dtype == F64) { const double *s = (const double *)r.data.data(); for (size_t i = 0; i < n; ++i) dst[i] = (real)s[i]; }
		else throw Exception(Exception::Type::ArgumentError, "state array has wrong dtype");
	}

	// ---- command line
	struct Options
	{
		std::string taskFile, resourcePath, out, loadState, mode = "run";
		std::vector<int> dumpSteps;
		std::string resumeFrom; // --resume <file>: start from a checkpoint file on disk (src/SpinParser.cpp:131-135), e.g. one written by an earlier process (a copy:
		                        // the task-file parser of a fresh start removes <task>.checkpoint, src/TaskFileParser.cpp:74-100)
		int maxSteps = -1, startStep = 0, threads = 0, timeStride = 1, timeRepeat = 1, timeWarmup = 1, timeOffset = 0, resumeAfter = -1;
		bool measure = true, verbose = false, dumpLattice = true, timeCompact = false;
	};
	static Options opt;

	inline std::vector<int> parseIntList(const std::string &s) { std::vector<int> v; std::stringstream ss(s); std::string t; while (std::getline(ss, t, ',')) if (!t.empty()) v.push_back(std::stoi(t)); return v; }

	// ---- accessors for the three cores' state (reference layout: SU2 {SS,DD}, XYZ {XX,YY,ZZ,DD}, TRI {one array})
	struct StateView { std::string core; int nChannelArrays; real *v4[4]; size_t v4size; real *v2; int v2size; real *cutoff; };
	inline StateView view(const std::string &core, EffectiveAction *a)
	{
		StateView s; s.core = core; s.cutoff = &a->cutoff;
		if (core == "SU2") { auto *e = static_cast<SU2EffectiveAction *>(a); s.nChannelArrays = 2; s.v4[0] = e->vertexTwoParticle->_dataSS; s.v4[1] = e->vertexTwoParticle->_dataDD; s.v4size = e->vertexTwoParticle->size; s.v2 = e->vertexSingleParticle->_data; s.v2size = e->vertexSingleParticle->size; }
		else if (core == "XYZ") { auto *e = static_cast<XYZEffectiveAction *>(a); s.nChannelArrays = 4; s.v4[0] = e->vertexTwoParticle->_dataXX; s.v4[1] = e->vertexTwoParticle->_dataYY; s.v4[2] = e->vertexTwoParticle->_dataZZ; s.v4[3] = e->vertexTwoParticle->_dataDD; s.v4size = e->vertexTwoParticle->size; s.v2 = e->vertexSingleParticle->_data; s.v2size = e->vertexSingleParticle->size; }
		else { auto *e = static_cast<TRIEffectiveAction *>(a); s.nChannelArrays = 1; s.v4[0] = e->vertexTwoParticle->_data; s.v4size = e->vertexTwoParticle->size; s.v2 = e->vertexSingleParticle->_data; s.v2size = e->vertexSingleParticle->size; }
		return s;
	}
	inline void dumpState(Writer &w, const std::string &prefix, const StateView &s)
	{
		w.scalar(prefix + "/cutoff", (double)*s.cutoff);
		w.reals(prefix + "/v2", s.v2, { (uint64_t)s.v2size });
		for (int c = 0; c < s.nChannelArrays; ++c) w.reals(prefix + "/v4_" + std::to_string(c), s.v4[c], { (uint64_t)s.v4size });
	}

	inline void dumpLattice(Writer &w)
	{
		const Lattice &l = FrgCommon::lattice();
		int L = l.size, N = l.end() - l.begin(), nb = (int)l._basis.size();
		w.scalar("lattice/size", L); w.scalar("lattice/dataSize", N); w.scalar("lattice/nBasis", nb);
		auto desc = [&](const LatticeSiteDescriptor *d, int n, const std::string &name)
		{
			std::vector<int> rid(n), perm(3 * n);
			for (int i = 0; i < n; ++i) { rid[i] = d[i].rid; for (int k = 0; k < 3; ++k) perm[3 * i + k] = static_cast<int>(d[i].spinPermutation[k]); }
			w.ints(name + "_rid", rid); w.ints(name + "_perm", perm, { (uint64_t)n, 3 });
		};
		desc(l.getSites(), L, "lattice/sites");
		desc(l.getInvertedSites(), L, "lattice/invertedSites");
		// overlap CSR (src/Lattice.hpp:46-150)
		std::vector<int> offs(L + 1, 0), r1, r2, p1, p2;
		for (int r = 0; r < L; ++r)
		{
			const LatticeOverlap &o = l.getOverlap(r);
			offs[r + 1] = offs[r] + o.size;
			for (int i = 0; i < o.size; ++i)
			{
				r1.push_back(o.rid1[i]); r2.push_back(o.rid2[i]);
				p1.push_back(static_cast<int>(o.transformedX1[i])); p1.push_back(static_cast<int>(o.transformedY1[i])); p1.push_back(static_cast<int>(o.transformedZ1[i]));
				p2.push_back(static_cast<int>(o.transformedX2[i])); p2.push_back(static_cast<int>(o.transformedY2[i])); p2.push_back(static_cast<int>(o.transformedZ2[i]));
			}
		}
		w.ints("lattice/overlap_offsets", offs); w.ints("lattice/overlap_rid1", r1); w.ints("lattice/overlap_rid2", r2);
		w.ints("lattice/overlap_perm1", p1, { (uint64_t)r1.size(), 3 }); w.ints("lattice/overlap_perm2", p2, { (uint64_t)r2.size(), 3 });
		// basis ids and, for every basis site b, its range list with the symmetry reduction of (b, j) and (j, b)
		std::vector<int> basis; for (auto b = l.getBasis(); b != l.end(); ++b) basis.push_back(b - l.begin());
		w.ints("lattice/basis", basis);
		for (int b = 0; b < nb; ++b)
		{
			std::vector<int> ids, fr, fp, ir, ip;
			LatticeIterator bi(basis[b]);
			for (auto j = l.getRange(b); j != l.end(); ++j)
			{
				ids.push_back(j - l.begin());
				SpinComponent x = SpinComponent::X, y = SpinComponent::Y, z = SpinComponent::Z;
				fr.push_back(l.symmetryTransform(bi, j, x, y, z)); fp.push_back(static_cast<int>(x)); fp.push_back(static_cast<int>(y)); fp.push_back(static_cast<int>(z));
				x = SpinComponent::X; y = SpinComponent::Y; z = SpinComponent::Z;
				ir.push_back(l.symmetryTransform(j, bi, x, y, z)); ip.push_back(static_cast<int>(x)); ip.push_back(static_cast<int>(y)); ip.push_back(static_cast<int>(z));
			}
			std::string p = "lattice/range" + std::to_string(b);
			w.ints(p + "_ids", ids); w.ints(p + "_fwd_rid", fr); w.ints(p + "_fwd_perm", fp, { (uint64_t)fr.size(), 3 });
			w.ints(p + "_inv_rid", ir); w.ints(p + "_inv_perm", ip, { (uint64_t)ir.size(), 3 });
		}
		// geometry (for the .obs meta groups)
		std::vector<double> bl; for (auto &a : l._bravaisLattice) { bl.push_back(a.x); bl.push_back(a.y); bl.push_back(a.z); }
		w.doubles("lattice/bravais", bl, { (uint64_t)l._bravaisLattice.size(), 3 });
		std::vector<double> pos; std::vector<int> par;
		for (int i = 0; i < N; ++i) { auto p = l.getSitePosition(LatticeIterator(i)); pos.push_back(p.x); pos.push_back(p.y); pos.push_back(p.z); auto t = l.getSiteParameters(LatticeIterator(i)); par.push_back(std::get<0>(t)); par.push_back(std::get<1>(t)); par.push_back(std::get<2>(t)); par.push_back(std::get<3>(t)); }
		w.doubles("lattice/positions", pos, { (uint64_t)N, 3 }); w.ints("lattice/parameters", par, { (uint64_t)N, 4 });
	}

	inline void dumpMeshes(Writer &w)
	{
		const FrequencyDiscretization &fd = FrgCommon::frequency();
		w.reals("frequency", fd._data, { (uint64_t)fd.size });
		std::vector<double> c; for (auto i = FrgCommon::cutoff().begin(); i != FrgCommon::cutoff().end(); ++i) c.push_back((double)*i);
		w.doubles("cutoff", c);
	}

	// every dataset the reference wrote through the HDF5 API (obs / checkpoint files), with the cutoff attribute of its group
	inline void dumpH5(Writer &w)
	{
		std::function<void(const std::string &, const std::shared_ptr<h5shim::Node> &)> walk = [&](const std::string &path, const std::shared_ptr<h5shim::Node> &n)
		{
			auto a = n->attributes.find("cutoff");
			if (a != n->attributes.end()) w.reals("h5" + path + "@cutoff", (const real *)a->second.data(), { 1 });
			if (!n->isGroup)
			{
				std::vector<uint64_t> dims(n->dims.begin(), n->dims.end());
				size_t per = n->elemSize / sizeof(real);
				if (per > 1) dims.push_back(per);
				w.reals("h5" + path, (const real *)n->data.data(), dims);
			}
			for (auto &c : n->children) walk(path + "/" + c.first, c.second);
		};
		for (auto &f : h5shim::files()) { std::string tag = boost::filesystem::path(f.first).extension().string(); walk("/" + (tag.empty() ? std::string("file") : tag.substr(1)), f.second); }
	}
}

// ===================================================================== CommandLineOptions replacement bodies
CommandLineOptions::CommandLineOptions(int argc, char **argv)
{
	using harness::opt;
	_help = false; _verbose = false; _checkpointTime = 1 << 30; _forceRestart = true; _deferMeasurements = false; _debugLattice = false;
	for (int i = 1; i < argc; ++i)
	{
		std::string a = argv[i];
		auto next = [&]() -> std::string { if (i + 1 >= argc) throw Exception(Exception::Type::ArgumentError, "missing value for " + a); return argv[++i]; };
		if (a == "-r" || a == "--resourcePath") _resourcePath = next();
		else if (a == "-v" || a == "--verbose") _verbose = true;
		else if (a == "-d" || a == "--defer") _deferMeasurements = true;
		else if (a == "--out") opt.out = next();
		else if (a == "--mode") opt.mode = next();
		else if (a == "--dump-steps") opt.dumpSteps = harness::parseIntList(next());
		else if (a == "--max-steps") opt.maxSteps = std::stoi(next());
		else if (a == "--start-step") opt.startStep = std::stoi(next());
		else if (a == "--load-state") opt.loadState = next();
		else if (a == "--threads") opt.threads = std::stoi(next());
		else if (a == "--time-stride") opt.timeStride = std::stoi(next());
		else if (a == "--time-repeat") opt.timeRepeat = std::stoi(next());
		else if (a == "--time-warmup") opt.timeWarmup = std::stoi(next());
		else if (a == "--time-offset") opt.timeOffset = std::stoi(next());
		else if (a == "--time-compact") opt.timeCompact = true;
		else if (a == "--resume-after") opt.resumeAfter = std::stoi(next());
		else if (a == "--resume") opt.resumeFrom = next();
		else if (a == "--checkpoint-time") _checkpointTime = std::stoi(next());
		else if (a == "--no-measure") opt.measure = false;
		else if (a == "--no-lattice") opt.dumpLattice = false;
		else if (a.size() && a[0] == '-') throw Exception(Exception::Type::ArgumentError, "unknown option " + a);
		else _taskFile = a;
	}
	opt.taskFile = _taskFile; opt.resourcePath = _resourcePath; opt.verbose = _verbose;
}
bool CommandLineOptions::help() const { return _help; }
bool CommandLineOptions::verbose() const { return _verbose; }
int CommandLineOptions::checkpointTime() const { return _checkpointTime; }
bool CommandLineOptions::forceRestart() const { return _forceRestart; }
bool CommandLineOptions::deferMeasurements() const { return _deferMeasurements; }
bool CommandLineOptions::debugLattice() const { return _debugLattice; }
std::string CommandLineOptions::taskFile() const { return _taskFile; }
std::string CommandLineOptions::resourcePath() const { return _resourcePath; }

// ===================================================================== SpinParser replacement bodies
SpinParser *SpinParser::_spinParserInstance = nullptr;
SpinParser *SpinParser::spinParser() { if (_spinParserInstance == nullptr) _spinParserInstance = new SpinParser; return _spinParserInstance; }
FrgCore *SpinParser::getFrgCore() const { return _frgCore; }
SpinParser::SpinParser() { _isMasterRank = !getenv("PFFRG_RANK") || atoi(getenv("PFFRG_RANK")) == 0; _commandLineOptions = nullptr; _taskFileParser = nullptr; _loadManager = HMP::newLoadManager(); _frgCore = nullptr; }
SpinParser::~SpinParser() { delete _commandLineOptions; delete _frgCore; }
bool SpinParser::isMasterRank() const { return _isMasterRank; }
ComputationStatus SpinParser::getComputationStatus() const { return _computationStatus; }
Fileset SpinParser::getFileset() const { return _fileset; }
CommandLineOptions *SpinParser::getCommandLineOptions() const { return _commandLineOptions; }
TaskFileParser *SpinParser::getTaskFileParser() const { return _taskFileParser; }
HMP::LoadManager *SpinParser::getLoadManager() const { return _loadManager; }
void SpinParser::runCore() {}
void SpinParser::writeCheckpoint() {}

int SpinParser::run(int argc, char **argv)
{
	using namespace harness;
	try
	{
		_commandLineOptions = new CommandLineOptions(argc, argv);
		Log::log << Log::setDisplayLogLevel(opt.verbose ? Log::LogLevel::Info : Log::LogLevel::Warning);
		if (opt.threads > 0) omp_set_num_threads(opt.threads);

		_fileset.taskFile = opt.taskFile;
		_fileset.obsFile = boost::filesystem::path(_fileset.taskFile).replace_extension("obs").string();
		_fileset.dataFile = boost::filesystem::path(_fileset.taskFile).replace_extension("data").string();
		_fileset.checkpointFile = boost::filesystem::path(_fileset.taskFile).replace_extension("checkpoint").string();

		// the reference's own task-file parser builds meshes, lattice, spin model and the FRG core (src/TaskFileParser.cpp:21-347)
		auto t0 = std::chrono::steady_clock::now();
		_taskFileParser = new TaskFileParser(_fileset.taskFile, FrgCommon::_frequency, FrgCommon::_cutoff, FrgCommon::_lattice, _frgCore, _computationStatus);
		double setupSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		_computationStatus.statusIdentifier = ComputationStatus::Identifier::New;

		std::string core = "SU2";
		if (dynamic_cast<XYZFrgCore *>(_frgCore)) core = "XYZ";
		else if (dynamic_cast<TRIFrgCore *>(_frgCore)) core = "TRI";

		Writer w; if (!opt.out.empty() && !w.open(opt.out)) throw Exception(Exception::Type::IOError, "cannot open output " + opt.out);
		w.raw("core", U8, { (uint64_t)core.size() }, core.data(), core.size());
		w.scalar("realBytes", (double)sizeof(real));
		w.scalar("setupSeconds", setupSeconds);
		if (core == "SU2") { w.scalar("spinLength", (double)static_cast<SU2FrgCore *>(_frgCore)->spinLength); w.scalar("normalization", (double)static_cast<SU2FrgCore *>(_frgCore)->normalization); }
		else if (core == "XYZ") w.scalar("normalization", (double)static_cast<XYZFrgCore *>(_frgCore)->normalization);
		else w.scalar("normalization", (double)static_cast<TRIFrgCore *>(_frgCore)->normalization);
		dumpMeshes(w);
		if (opt.dumpLattice) dumpLattice(w);

		StateView state = view(core, _frgCore->_flowingFunctional);
		StateView flow = view(core, _frgCore->_flow);
		std::vector<double> cutoffs; for (auto i = FrgCommon::cutoff().begin(); i != FrgCommon::cutoff().end(); ++i) cutoffs.push_back((double)*i);
		dumpState(w, "initial", state);

		// optional state injection (synthetic benchmark states, round-trip tests)
		int step = opt.startStep;
		if (!opt.loadState.empty())
		{
			auto recs = readPfd(opt.loadState);
			loadReals(recs.at("v2"), state.v2, state.v2size);
			for (int c = 0; c < state.nChannelArrays; ++c) loadReals(recs.at("v4_" + std::to_string(c)), state.v4[c], state.v4size);
		}
		CutoffIterator cutoff = FrgCommon::cutoff().begin();
		for (int i = 0; i < step; ++i) ++cutoff;
		*state.cutoff = *cutoff;
		if (!opt.resumeFrom.empty())
		{
			// src/SpinParser.cpp:131-135: continue from the checkpoint file (written by an earlier process through the same HDF5 calls)
			if (!_frgCore->_flowingFunctional->readCheckpoint(opt.resumeFrom)) throw Exception(Exception::Type::IOError, "no checkpoint to resume from");
			cutoff = FrgCommon::cutoff().find(_frgCore->_flowingFunctional->cutoff);
			step = 0; for (auto i = FrgCommon::cutoff().begin(); i != cutoff; ++i) ++step;
			w.scalar("resumedFromStep", step);
		}

		if (opt.mode == "time")
		{
			// CPU baseline: the reference's per-item calculators (src/SU2/SU2FrgCore.cpp:139-431 and the XYZ/TRI equivalents) under the
			// same `omp parallel for schedule(guided)` the reference's LoadManager uses (src/lib/LoadManager.hpp:551-557), on a strided item sample.
			int nw = FrgCommon::frequency().size, nf = nw * nw * (nw + 1) / 2;
			*flow.cutoff = *state.cutoff;
			auto v2item = [&](int i) { if (core == "SU2") static_cast<SU2FrgCore *>(_frgCore)->_calculateVertexSingleParticle(i); else if (core == "XYZ") static_cast<XYZFrgCore *>(_frgCore)->_calculateVertexSingleParticle(i); else static_cast<TRIFrgCore *>(_frgCore)->_calculateVertexSingleParticle(i); };
			auto v4item = [&](int i) { if (core == "SU2") static_cast<SU2FrgCore *>(_frgCore)->_calculateVertexTwoParticle(i); else if (core == "XYZ") static_cast<XYZFrgCore *>(_frgCore)->_calculateVertexTwoParticle(i); else static_cast<TRIFrgCore *>(_frgCore)->_calculateVertexTwoParticle(i); };
			std::vector<int> items; for (int i = opt.timeOffset; i < nf; i += opt.timeStride) items.push_back(i);
			std::vector<double> times;
			for (int rep = 0; rep < opt.timeWarmup + opt.timeRepeat; ++rep)
			{
				auto tic = std::chrono::steady_clock::now();
				#pragma omp parallel for schedule(static)
				for (int i = 0; i < nw; ++i) v2item(i);
				#pragma omp parallel for schedule(guided)
				for (int k = 0; k < (int)items.size(); ++k) v4item(items[k]);
				double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - tic).count();
				if (rep >= opt.timeWarmup) times.push_back(dt);
			}
			w.doubles("time/seconds", times);
			w.scalar("time/items", (double)items.size()); w.scalar("time/itemsTotal", (double)nf); w.scalar("time/threads", (double)omp_get_max_threads());
			w.ints("time/itemIds", items);
			if (opt.timeCompact)
			{
				// only the sampled items' rows (size-parity tests at the benchmark sizes: the full arrays are hundreds of MB)
				const size_t per = flow.v4size / (size_t)nf;
				w.reals("time/flow/v2", flow.v2, { (uint64_t)flow.v2size });
				for (int c = 0; c < flow.nChannelArrays; ++c)
				{
					std::vector<real> rows(items.size() * per);
					for (size_t k = 0; k < items.size(); ++k) std::copy(flow.v4[c] + (size_t)items[k] * per, flow.v4[c] + (size_t)(items[k] + 1) * per, rows.begin() + k * per);
					w.reals("time/flowItems/v4_" + std::to_string(c), rows.data(), { (uint64_t)items.size(), (uint64_t)per });
				}
			}
			else dumpState(w, "time/flow", flow);
			printf("{\"core\": \"%s\", \"items\": %d, \"items_total\": %d, \"threads\": %d, \"seconds\": [", core.c_str(), (int)items.size(), nf, omp_get_max_threads());
			for (size_t i = 0; i < times.size(); ++i) printf("%s%.6f", i ? ", " : "", times[i]);
			printf("]}\n");
			w.close();
			return 0;
		}

		// Euler loop, src/SpinParser.cpp:141-172 (restated here: SpinParser.cpp itself is not compiled). --resume-after N: a checkpoint is
		// written after N steps (SpinParser::writeCheckpoint, :224-233, its vertex part); after the run the checkpoint is read back
		// (:131-135) and the flow repeated from there without measurements: "resumed/final" must equal "final".
		auto wants = [&](int s) { return std::find(opt.dumpSteps.begin(), opt.dumpSteps.end(), s) != opt.dumpSteps.end(); };
		std::vector<double> stepSeconds;
		_computationStatus.checkpointTime = Timestamp::time();
		auto runLoop = [&](CutoffIterator cutoff, int step, bool measure, bool first)
		{
			int done = 0;
			while (cutoff != FrgCommon::cutoff().last())
			{
				if (opt.maxSteps >= 0 && done >= opt.maxSteps) break;
				if (first && wants(step)) dumpState(w, "step" + std::to_string(step) + "/state", state);
				auto tic = std::chrono::steady_clock::now();
				_frgCore->computeStep();
				if (first) stepSeconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - tic).count());
				if (measure) _frgCore->takeMeasurements();
				if (first && wants(step)) dumpState(w, "step" + std::to_string(step) + "/flow", flow);
				if (_frgCore->_flow->isDiverged()) { if (first) w.scalar("divergedAtStep", step); break; }
				++cutoff; ++step; ++done;
				// --resume-after: make the periodic checkpoint of the driver loop due after this step (the last one was "long ago")
				if (first && done == opt.resumeAfter) _computationStatus.checkpointTime = Timestamp::time() + boost::posix_time::time_duration(-24 * 365 * 100, 0, 0, 0);
				_frgCore->finalizeStep(*cutoff);
				if (Timestamp::isOlder(_computationStatus.checkpointTime, _commandLineOptions->checkpointTime()))
				{
					_computationStatus.checkpointTime = Timestamp::time();
					_frgCore->_flowingFunctional->writeCheckpoint(_fileset.checkpointFile);
					if (first) w.scalar("checkpointAtStep", step);
				}
			}
			if (measure) _frgCore->takeMeasurements();
			return step;
		};
		step = runLoop(cutoff, step, opt.measure, true);
		dumpState(w, "final", state);
		w.scalar("finalStep", step);
		w.doubles("stepSeconds", stepSeconds);
		if (opt.resumeAfter >= 0)
		{
			if (!_frgCore->_flowingFunctional->readCheckpoint(_fileset.checkpointFile)) throw Exception(Exception::Type::IOError, "no checkpoint to resume from");
			CutoffIterator resumed = FrgCommon::cutoff().find(_frgCore->_flowingFunctional->cutoff);
			int resumedStep = 0; for (auto i = FrgCommon::cutoff().begin(); i != resumed; ++i) ++resumedStep;
			w.scalar("resumed/fromStep", resumedStep);
			const int last = runLoop(resumed, resumedStep, false, false);
			dumpState(w, "resumed/final", state);
			w.scalar("resumed/finalStep", last);
		}
		// post-processing stage of deferred measurements, src/SpinParser.cpp:199-214: every state FrgCore::takeMeasurements appended to
		// the data file is read back and measured
		bool postprocessing = _commandLineOptions->deferMeasurements();
		for (auto m : _frgCore->_measurements) if (m->isDeferred()) postprocessing = true;
		if (postprocessing && opt.measure)
		{
			_computationStatus.statusIdentifier = ComputationStatus::Identifier::Postprocessing;
			int n = 0;
			while (_frgCore->_flowingFunctional->readCheckpoint(_fileset.dataFile, n++)) _frgCore->takeMeasurements();
			w.scalar("postprocessedStates", n - 1);
		}
		dumpH5(w);
		w.close();
	}
	catch (std::exception &e)
	{
		fprintf(stderr, "ref_harness: caught exception: %s\n", e.what());
		return 1;
	}
	return 0;
}

int main(int argc, char **argv)
{
	return SpinParser::spinParser()->run(argc, argv);
}
