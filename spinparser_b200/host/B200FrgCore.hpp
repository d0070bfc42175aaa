// B200FrgCore.hpp -- SpinParser flow cores whose computeStep() / finalizeStep() run on a B200 through libpffrg.
//
// This header is the reference-side half of the drop-in boundary (the other half is include/pffrg.h). It is meant to
// be compiled INSIDE the SpinParser source tree, next to src/FrgCoreFactory.cpp: it includes the
// reference's own headers and derives from the reference's own core classes, so that everything else in SpinParser --
// task file parsing, lattice construction, measurements (which static_cast the core and the effective action to the
// concrete SU2/XYZ/TRI types, e.g. src/SU2/SU2MeasurementCorrelation.cpp:81-84), checkpointing, the Euler loop of
// src/SpinParser.cpp:141-172 -- keeps working unchanged.
//
//   B200FrgCore<SU2FrgCore>, B200FrgCore<XYZFrgCore>, B200FrgCore<TRIFrgCore>
//
// inherit the stock constructor (core options, initial condition, host arrays in reference layout) and replace the two
// virtuals of src/FrgCore.hpp:77,86. The host arrays remain the interface to the rest of SpinParser:
//   computeStep()   uploads the state if the host copy is newer, runs the flow on the GPU, and -- concurrently, on the
//                   host -- the load-managed measurement stacks exactly like src/SU2/SU2FrgCore.cpp:96-108 does;
//                   _flow receives the self-energy flow and the divergence signal (NaN), and the full vertex flow
//                   when syncFlow is set.
//   finalizeStep()  Euler update on the GPU (state stays resident in FP64). The reference's host arrays (page-locked in place) are
//                   refreshed only when somebody is about to read them (syncState=lazy, the default): host-side measurements and the
//                   deferred-measurement dump of FrgCore::takeMeasurements (src/FrgCore.hpp:36-68), the periodic and the final
//                   checkpoint of SpinParser::runCore (src/SpinParser.cpp:163-183), a diverged flow. syncState=always copies after
//                   every step. The host cutoff is always advanced (the driver loop reads it).
//
// Several GPUs: one process per GPU, as the reference runs one MPI rank per node. In an MPI build the ranks are the MPI ranks
// (the NCCL id travels by MPI_Bcast); without MPI the ranks are given by the environment (PFFRG_RANK, PFFRG_NRANKS, PFFRG_ID_FILE:
// rank 0 writes the id to that file, the others wait for it). Every rank runs the host program; the work items of every step are
// sharded and the updated slices exchanged over NVLink inside pffrg_finalize_step.
//
// The real type of the host arrays is `float` as in the reference; a build that redefines float as double (the FP64
// parity oracle of this repository) is handled by sizeof.
#pragma once

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "pffrg.h"
#ifndef DISABLE_MPI
#include <mpi.h>
#endif

#include "FrgCommon.hpp"
#include "SpinParser.hpp"
#include "lib/Exception.hpp"
#include "SU2/SU2FrgCore.hpp"
#include "SU2/SU2EffectiveAction.hpp"
#include "XYZ/XYZFrgCore.hpp"
#include "XYZ/XYZEffectiveAction.hpp"
#include "TRI/TRIFrgCore.hpp"
#include "TRI/TRIEffectiveAction.hpp"

namespace b200
{
	typedef float real;

	// host arrays of one effective action in reference layout
	struct HostArrays
	{
		real *v2 = nullptr; int v2Size = 0;
		real *v4[4] = { nullptr, nullptr, nullptr, nullptr }; size_t v4Size = 0; int nArrays = 0;
	};

	template <class RefCore> struct CoreTraits;
	template <> struct CoreTraits<SU2FrgCore>
	{
		static const int id = PFFRG_CORE_SU2;
		static double spinLength(const SU2FrgCore &c) { return c.spinLength; }
		static HostArrays arrays(EffectiveAction *a)
		{
			SU2EffectiveAction *e = static_cast<SU2EffectiveAction *>(a);
			HostArrays h; h.v2 = e->vertexSingleParticle->_data; h.v2Size = e->vertexSingleParticle->size;
			h.nArrays = 2; h.v4[0] = e->vertexTwoParticle->_dataSS; h.v4[1] = e->vertexTwoParticle->_dataDD; h.v4Size = e->vertexTwoParticle->size;
			return h;
		}
	};
	template <> struct CoreTraits<XYZFrgCore>
	{
		static const int id = PFFRG_CORE_XYZ;
		static double spinLength(const XYZFrgCore &) { return 0.5; }
		static HostArrays arrays(EffectiveAction *a)
		{
			XYZEffectiveAction *e = static_cast<XYZEffectiveAction *>(a);
			HostArrays h; h.v2 = e->vertexSingleParticle->_data; h.v2Size = e->vertexSingleParticle->size;
			h.nArrays = 4; h.v4[0] = e->vertexTwoParticle->_dataXX; h.v4[1] = e->vertexTwoParticle->_dataYY; h.v4[2] = e->vertexTwoParticle->_dataZZ; h.v4[3] = e->vertexTwoParticle->_dataDD;
			h.v4Size = e->vertexTwoParticle->size;
			return h;
		}
	};
	template <> struct CoreTraits<TRIFrgCore>
	{
		static const int id = PFFRG_CORE_TRI;
		static double spinLength(const TRIFrgCore &) { return 0.5; }
		static HostArrays arrays(EffectiveAction *a)
		{
			TRIEffectiveAction *e = static_cast<TRIEffectiveAction *>(a);
			HostArrays h; h.v2 = e->vertexSingleParticle->_data; h.v2Size = e->vertexSingleParticle->size;
			h.nArrays = 1; h.v4[0] = e->vertexTwoParticle->_data; h.v4Size = e->vertexTwoParticle->size;
			return h;
		}
	};

	// The tables the hot path reads from FrgCommon::lattice() / FrgCommon::frequency(), flattened for pffrg_desc.
	struct ProblemTables
	{
		std::vector<double> frequencies;
		std::vector<int32_t> sitesRid, sitesPerm, invertedRid, invertedPerm;
		std::vector<int32_t> overlapOffsets, overlapRid1, overlapRid2, overlapPerm1, overlapPerm2;
		std::vector<int32_t> rangeFwd, rangeInv;

		ProblemTables()
		{
			const FrequencyDiscretization &f = FrgCommon::frequency();
			for (int i = 0; i < f.size; ++i) frequencies.push_back((double)f._data[i]);
			const Lattice &l = FrgCommon::lattice();
			const int L = l.size;
			for (int j = 0; j < L; ++j)
			{
				sitesRid.push_back(l.getSites()[j].rid); invertedRid.push_back(l.getInvertedSites()[j].rid);
				for (int k = 0; k < 3; ++k)
				{
					sitesPerm.push_back(static_cast<int>(l.getSites()[j].spinPermutation[k]));
					invertedPerm.push_back(static_cast<int>(l.getInvertedSites()[j].spinPermutation[k]));
				}
			}
			overlapOffsets.push_back(0);
			for (int r = 0; r < L; ++r)
			{
				const LatticeOverlap &o = l.getOverlap(r);
				overlapOffsets.push_back(overlapOffsets.back() + o.size);
				for (int i = 0; i < o.size; ++i)
				{
					overlapRid1.push_back(o.rid1[i]); overlapRid2.push_back(o.rid2[i]);
					overlapPerm1.push_back(static_cast<int>(o.transformedX1[i])); overlapPerm1.push_back(static_cast<int>(o.transformedY1[i])); overlapPerm1.push_back(static_cast<int>(o.transformedZ1[i]));
					overlapPerm2.push_back(static_cast<int>(o.transformedX2[i])); overlapPerm2.push_back(static_cast<int>(o.transformedY2[i])); overlapPerm2.push_back(static_cast<int>(o.transformedZ2[i]));
				}
			}
			for (auto j = l.getRange(0); j != l.end(); ++j)
			{
				rangeFwd.push_back(l.symmetryTransform(l.zero(), j));
				rangeInv.push_back(l.symmetryTransform(j, l.zero()));
			}
		}

		pffrg_desc descriptor(int core, double spinLength, int device) const
		{
			pffrg_desc d;
			d.abi_version = PFFRG_ABI_VERSION; d.core = core;
			d.n_frequencies = (int32_t)frequencies.size(); d.frequencies = frequencies.data();
			d.n_sites = (int32_t)sitesRid.size();
			d.sites_rid = sitesRid.data(); d.sites_perm = sitesPerm.data(); d.inverted_rid = invertedRid.data(); d.inverted_perm = invertedPerm.data();
			d.overlap_offsets = overlapOffsets.data(); d.overlap_rid1 = overlapRid1.data(); d.overlap_rid2 = overlapRid2.data();
			d.overlap_perm1 = overlapPerm1.data(); d.overlap_perm2 = overlapPerm2.data();
			d.n_range = (int32_t)rangeFwd.size(); d.range_fwd_rid = rangeFwd.data(); d.range_inv_rid = rangeInv.data();
			d.spin_length = spinLength; d.device = device;
			return d;
		}
	};

	// options consumed by the adapter (removed before the stock constructor sees the option list, which rejects unknown keys:
	// src/SU2/SU2FrgCore.cpp:23-28)
	struct AdapterOptions
	{
		int device = -1;        // -1: the local rank of a multi-GPU run, else 0
		bool syncState = false; // true ("always"): copy the updated state into the reference's host arrays after every finalizeStep;
		                        // false ("lazy", default): only when the host arrays are about to be read
		bool syncFlow = false;  // copy the full vertex flow into _flow after every computeStep
		bool deviceMeasurement = true; // correlation measurements on the GPU (B200MeasurementCorrelation) instead of the reference's host code
		static AdapterOptions extract(std::map<std::string, std::string> &options)
		{
			AdapterOptions a;
			if (const char *e = getenv("SPINPARSER_B200_DEVICE")) a.device = atoi(e);
			auto always = [](const std::string &v) { return v == "true" || v == "1" || v == "always"; };
			if (const char *e = getenv("SPINPARSER_B200_SYNC_STATE")) a.syncState = always(e);
			if (const char *e = getenv("SPINPARSER_B200_SYNC_FLOW")) a.syncFlow = atoi(e) != 0;
			if (const char *e = getenv("SPINPARSER_B200_MEASUREMENT")) a.deviceMeasurement = std::string(e) != "host";
			auto take = [&](const char *key, std::string &out) { auto it = options.find(key); if (it == options.end()) return false; out = it->second; options.erase(it); return true; };
			std::string v;
			if (take("device", v)) a.device = std::stoi(v);
			if (take("syncState", v))
			{
				if (!always(v) && v != "false" && v != "0" && v != "lazy") throw Exception(Exception::Type::InitializationError, "Unknown syncState policy '" + v + "'.");
				a.syncState = always(v);
			}
			if (take("syncFlow", v)) a.syncFlow = (v == "true" || v == "1");
			if (take("measurement", v))
			{
				if (v != "host" && v != "device") throw Exception(Exception::Type::InitializationError, "Unknown measurement backend '" + v + "'.");
				a.deviceMeasurement = v == "device";
			}
			take("backend", v);
			return a;
		}
		static std::map<std::string, std::string> strip(std::map<std::string, std::string> options) { extract(options); return options; }
	};

	// measurements that read the DEVICE state (B200MeasurementCorrelation): the host arrays need not be refreshed for them
	struct DeviceMeasurement { virtual ~DeviceMeasurement() {} };

	template <class RefCore>
	class B200FrgCore : public RefCore
	{
	public:
		B200FrgCore(const SpinModel &spinModel, const std::vector<Measurement *> &measurements, const std::map<std::string, std::string> &options)
			: RefCore(spinModel, measurements, AdapterOptions::strip(options)), _handle(nullptr), _deviceCutoff(NAN), _deviceCurrent(false)
		{
			std::map<std::string, std::string> copy = options;
			_options = AdapterOptions::extract(copy);
			int rank = 0, nRanks = 1;
			rankLayout(rank, nRanks);
			if (_options.device < 0) { const int devices = pffrg_device_count(); _options.device = devices > 0 ? rank % devices : 0; }
			const ProblemTables tables;
			const pffrg_desc desc = tables.descriptor(CoreTraits<RefCore>::id, CoreTraits<RefCore>::spinLength(*this), _options.device);
			check(pffrg_create(&desc, &_handle), "pffrg_create");
			if (nRanks > 1) joinRanks(rank, nRanks);
			_state = CoreTraits<RefCore>::arrays(this->_flowingFunctional);
			_flowArrays = CoreTraits<RefCore>::arrays(this->_flow);
			if ((int64_t)_state.v4Size != pffrg_vertex_array_length(_handle) || _state.nArrays != pffrg_num_vertex_arrays(_handle))
				throw Exception(Exception::Type::InternalError, "B200FrgCore: vertex layout of libpffrg does not match the host arrays");
			// page-lock the reference's own arrays in place: transfers run at full PCIe speed without an extra host copy
			for (int c = 0; c < _state.nArrays; ++c) { pffrg_host_register(_state.v4[c], _state.v4Size * sizeof(real)); if (_options.syncFlow) pffrg_host_register(_flowArrays.v4[c], _flowArrays.v4Size * sizeof(real)); }
		}

		~B200FrgCore()
		{
			for (int c = 0; c < _state.nArrays; ++c) { pffrg_host_unregister(_state.v4[c]); if (_options.syncFlow) pffrg_host_unregister(_flowArrays.v4[c]); }
			pffrg_destroy(_handle);
		}

		// FrgCore::computeStep, src/FrgCore.hpp:77; reference implementation src/SU2/SU2FrgCore.cpp:89-109
		void computeStep() override
		{
			this->_flow->cutoff = this->_flowingFunctional->cutoff;
			// the host arrays are the truth whenever somebody else wrote them (construction, readCheckpoint): detected by the cutoff and
			// by a fingerprint of the arrays taken when host and device were last known to agree
			if (hostWasModified()) uploadState();

			// who reads the host arrays in the takeMeasurements() that follows this call (src/FrgCore.hpp:36-68)?
			const bool deferAll = SpinParser::spinParser()->getCommandLineOptions()->deferMeasurements();
			const float cutoff = this->_flowingFunctional->cutoff;
			bool hostReaders = false;
			std::vector<HMP::StackIdentifier> managed;
			for (auto m : this->_measurements)
			{
				const bool inRange = cutoff <= m->maxCutoff() && cutoff >= m->minCutoff();
				if (deferAll || m->isDeferred()) hostReaders = true; // the state is appended to the data file after every step
				else if (inRange && !isDeviceMeasurement(m)) hostReaders = true;
				if (m->isLoadManaged()) { auto s = m->getLoadManagedStacks(); managed.insert(managed.end(), s.begin(), s.end()); }
			}
			if (hostReaders || !managed.empty()) ensureHostCurrent();

			// load-managed measurements read the host state at this cutoff. Without MPI they run on a helper thread while the GPU computes
			// the flow; in an MPI build LoadManager::calculate talks to the other ranks and must stay on the thread that initialised MPI
			// (MPI_Init: thread level SINGLE, src/main.cpp:10), so it runs first.
			std::exception_ptr measurementError;
			std::thread measurementThread;
			if (!managed.empty())
			{
#ifndef DISABLE_MPI
				SpinParser::spinParser()->getLoadManager()->calculate(managed.data(), int(managed.size()));
#else
				measurementThread = std::thread([&] {
					try { SpinParser::spinParser()->getLoadManager()->calculate(managed.data(), int(managed.size())); }
					catch (...) { measurementError = std::current_exception(); }
				});
#endif
			}

			int diverged = 0;
			const int rc = pffrg_compute_step(_handle, &diverged);
			if (measurementThread.joinable()) measurementThread.join();
			check(rc, "pffrg_compute_step");
			if (measurementError) std::rethrow_exception(measurementError);
			// a diverged flow ends the driver loop: the final measurements and the last checkpoint read the host arrays
			if (diverged) ensureHostCurrent();

			void *v4[4] = { _flowArrays.v4[0], _flowArrays.v4[1], _flowArrays.v4[2], _flowArrays.v4[3] };
			check(pffrg_get_flow(_handle, _flowArrays.v2, _options.syncFlow ? v4 : nullptr, dtype()), "pffrg_get_flow");
			// divergence is signalled through NaN in _flow (EffectiveAction::isDiverged, src/SpinParser.cpp:151)
			if (diverged) _flowArrays.v2[0] = NAN;
		}

		// FrgCore::finalizeStep, src/FrgCore.hpp:86; reference implementation src/SU2/SU2FrgCore.cpp:111-137
		void finalizeStep(float newCutoff) override
		{
			check(pffrg_finalize_step(_handle, (double)newCutoff), "pffrg_finalize_step");
			_deviceCutoff = (double)newCutoff;
			_hostCurrent = false;
			this->_flowingFunctional->cutoff = newCutoff;
			// who reads the host arrays before the next computeStep()? The checkpoint of the driver loop when it is due
			// (src/SpinParser.cpp:163-168; two seconds of margin against the clock moving on between this test and the driver's), and
			// after the last step the final measurements and the last checkpoint (:172-183).
			const ComputationStatus status = SpinParser::spinParser()->getComputationStatus();
			const bool checkpointDue = (Timestamp::time() - status.checkpointTime).total_seconds() + 2 > SpinParser::spinParser()->getCommandLineOptions()->checkpointTime();
			const bool lastStep = newCutoff == *FrgCommon::cutoff().last();
			if (_options.syncState || checkpointDue || lastStep) ensureHostCurrent();
		}

		// copy the device state into the reference's host arrays
		void downloadState()
		{
			void *v4[4] = { _state.v4[0], _state.v4[1], _state.v4[2], _state.v4[3] };
			double cutoff = 0.0;
			check(pffrg_get_state(_handle, &cutoff, _state.v2, v4, dtype()), "pffrg_get_state");
			_hostCurrent = true;
			_hostFingerprint = fingerprint();
		}
		// ... unless they already hold the device state
		void ensureHostCurrent() { if (_deviceCurrent && !_hostCurrent) downloadState(); }
		bool hostIsCurrent() const { return _hostCurrent; }

		// correlations chi[c * L + rid] of `state` (the flowing functional), computed on the device; uploads the state first when
		// the host copy is the newer one (post-processing of deferred measurements reads checkpoints into the host arrays)
		void measureCorrelation(const EffectiveAction &state, std::vector<double> &chi)
		{
			if (&state != this->_flowingFunctional) throw Exception(Exception::Type::ArgumentError, "B200FrgCore::measureCorrelation: only the flowing functional can be measured");
			if (hostWasModified()) uploadState();
			chi.assign(size_t(pffrg_num_channels(_handle)) * FrgCommon::lattice().size, 0.0);
			check(pffrg_measure_correlation(_handle, chi.data()), "pffrg_measure_correlation");
		}

		// tell the core that the host arrays were modified without changing the cutoff
		void invalidateDeviceState() { _deviceCurrent = false; }

		pffrg_handle handle() const { return _handle; }

	private:
		static int dtype() { return sizeof(real) == 8 ? PFFRG_F64 : PFFRG_F32; }

		void uploadState()
		{
			const void *v4[4] = { _state.v4[0], _state.v4[1], _state.v4[2], _state.v4[3] };
			check(pffrg_set_state(_handle, (double)this->_flowingFunctional->cutoff, _state.v2, v4, dtype()), "pffrg_set_state");
			_deviceCutoff = (double)this->_flowingFunctional->cutoff;
			_deviceCurrent = true; _hostCurrent = true;
			_hostFingerprint = fingerprint();
		}

		// Did anybody write the host arrays since host and device last agreed (construction, EffectiveAction::readCheckpoint -- also of a
		// checkpoint at the very cutoff the device is at)? The cutoff and a fingerprint of the arrays decide: the whole self energy and
		// a strided sample of every vertex array. While the host copy is merely stale (lazy synchronisation) nobody has written it, the
		// fingerprint still matches, and the stale copy is NOT uploaded.
		bool hostWasModified() const
		{
			if (!_deviceCurrent) return true;
			if (!((double)this->_flowingFunctional->cutoff == _deviceCutoff)) return true;
			return fingerprint() != _hostFingerprint;
		}
		unsigned long long fingerprint() const
		{
			unsigned long long hash = 1469598103934665603ull;
			auto mix = [&](const real &x) { unsigned char b[sizeof(real)]; memcpy(b, &x, sizeof(real)); for (unsigned char c : b) { hash ^= c; hash *= 1099511628211ull; } };
			for (int i = 0; i < _state.v2Size; ++i) mix(_state.v2[i]);
			const size_t stride = _state.v4Size > 8192 ? _state.v4Size / 8192 : 1;
			for (int c = 0; c < _state.nArrays; ++c) { for (size_t i = 0; i < _state.v4Size; i += stride) mix(_state.v4[c][i]); mix(_state.v4[c][_state.v4Size - 1]); }
			return hash;
		}

		static bool isDeviceMeasurement(const Measurement *m) { return dynamic_cast<const DeviceMeasurement *>(m) != nullptr; }

		// ranks of a multi-GPU run: the MPI ranks, or (no MPI) PFFRG_RANK / PFFRG_NRANKS
		static void rankLayout(int &rank, int &nRanks)
		{
			rank = 0; nRanks = 1;
#ifndef DISABLE_MPI
			MPI_Comm_rank(MPI_COMM_WORLD, &rank); MPI_Comm_size(MPI_COMM_WORLD, &nRanks);
#else
			if (const char *e = getenv("PFFRG_NRANKS")) nRanks = std::max(1, atoi(e));
			if (const char *e = getenv("PFFRG_RANK")) rank = atoi(e);
			if (rank < 0 || rank >= nRanks) throw Exception(Exception::Type::InitializationError, "B200FrgCore: PFFRG_RANK outside [0, PFFRG_NRANKS)");
#endif
		}
		// rank 0 creates the communicator id and ships it to the other ranks; then every rank joins (pffrg_comm_init)
		void joinRanks(int rank, int nRanks)
		{
			unsigned char id[PFFRG_UNIQUE_ID_BYTES];
			if (rank == 0) check(pffrg_comm_unique_id(id), "pffrg_comm_unique_id");
#ifndef DISABLE_MPI
			MPI_Bcast(id, PFFRG_UNIQUE_ID_BYTES, MPI_BYTE, 0, MPI_COMM_WORLD);
#else
			const char *path = getenv("PFFRG_ID_FILE");
			if (!path) throw Exception(Exception::Type::InitializationError, "B200FrgCore: PFFRG_NRANKS > 1 needs PFFRG_ID_FILE (a path all ranks can reach)");
			if (rank == 0)
			{
				const std::string tmp = std::string(path) + ".tmp";
				FILE *f = fopen(tmp.c_str(), "wb");
				if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) throw Exception(Exception::Type::IOError, "B200FrgCore: cannot write " + tmp);
				fclose(f);
				if (rename(tmp.c_str(), path) != 0) throw Exception(Exception::Type::IOError, std::string("B200FrgCore: cannot create ") + path);
			}
			else
			{
				const auto t0 = std::chrono::steady_clock::now();
				for (;;)
				{
					if (FILE *f = fopen(path, "rb")) { const size_t n = fread(id, 1, sizeof id, f); fclose(f); if (n == sizeof id) break; }
					if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) throw Exception(Exception::Type::IOError, std::string("B200FrgCore: no communicator id appeared in ") + path);
					std::this_thread::sleep_for(std::chrono::milliseconds(20));
				}
			}
#endif
			check(pffrg_comm_init(_handle, id, rank, nRanks), "pffrg_comm_init");
		}

		static void check(int rc, const char *what)
		{
			if (rc != PFFRG_OK) throw Exception(rc == PFFRG_ERR_ARGUMENT ? Exception::Type::ArgumentError : Exception::Type::InternalError, std::string(what) + ": " + pffrg_last_error());
		}

		pffrg_handle _handle;
		AdapterOptions _options;
		HostArrays _state, _flowArrays;
		double _deviceCutoff;
		bool _deviceCurrent;
		bool _hostCurrent = true;                 // the host arrays hold the device state (or the device has none yet)
		unsigned long long _hostFingerprint = 0;  // of the host arrays when host and device last agreed
	};
}
