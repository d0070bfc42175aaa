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
//   finalizeStep()  Euler update on the GPU (state stays resident in FP64), then the updated state is copied back into
//                   the reference's arrays (page-locked in place) unless syncState is off.
//
// The real type of the host arrays is `float` as in the reference; a build that redefines float as double (the FP64
// parity oracle of this repository) is handled by sizeof.
#pragma once

#include <cmath>
#include <cstdlib>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "pffrg.h"

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
		int device = 0;
		bool syncState = true;  // copy the updated state into the reference's host arrays after every finalizeStep
		bool syncFlow = false;  // copy the full vertex flow into _flow after every computeStep
		bool deviceMeasurement = true; // correlation measurements on the GPU (B200MeasurementCorrelation) instead of the reference's host code
		static AdapterOptions extract(std::map<std::string, std::string> &options)
		{
			AdapterOptions a;
			if (const char *e = getenv("SPINPARSER_B200_DEVICE")) a.device = atoi(e);
			if (const char *e = getenv("SPINPARSER_B200_SYNC_STATE")) a.syncState = atoi(e) != 0;
			if (const char *e = getenv("SPINPARSER_B200_SYNC_FLOW")) a.syncFlow = atoi(e) != 0;
			if (const char *e = getenv("SPINPARSER_B200_MEASUREMENT")) a.deviceMeasurement = std::string(e) != "host";
			auto take = [&](const char *key, std::string &out) { auto it = options.find(key); if (it == options.end()) return false; out = it->second; options.erase(it); return true; };
			std::string v;
			if (take("device", v)) a.device = std::stoi(v);
			if (take("syncState", v)) a.syncState = (v == "true" || v == "1");
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

	template <class RefCore>
	class B200FrgCore : public RefCore
	{
	public:
		B200FrgCore(const SpinModel &spinModel, const std::vector<Measurement *> &measurements, const std::map<std::string, std::string> &options)
			: RefCore(spinModel, measurements, AdapterOptions::strip(options)), _handle(nullptr), _deviceCutoff(NAN), _deviceCurrent(false)
		{
			std::map<std::string, std::string> copy = options;
			_options = AdapterOptions::extract(copy);
			const ProblemTables tables;
			const pffrg_desc desc = tables.descriptor(CoreTraits<RefCore>::id, CoreTraits<RefCore>::spinLength(*this), _options.device);
			check(pffrg_create(&desc, &_handle), "pffrg_create");
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
			// the host arrays are the truth whenever somebody else wrote them (construction, readCheckpoint): detected by the cutoff
			if (!_deviceCurrent || !((double)this->_flowingFunctional->cutoff == _deviceCutoff)) uploadState();

			// load-managed measurements read the host state at this cutoff; they run on the host while the GPU computes the flow
			std::vector<HMP::StackIdentifier> managed;
			for (auto m : this->_measurements)
				if (m->isLoadManaged()) { auto s = m->getLoadManagedStacks(); managed.insert(managed.end(), s.begin(), s.end()); }
			if (!managed.empty() && !_options.syncState) throw Exception(Exception::Type::InitializationError, "B200FrgCore: load-managed measurements need syncState");
			std::exception_ptr measurementError;
			std::thread measurementThread;
			if (!managed.empty())
				measurementThread = std::thread([&] {
					try { SpinParser::spinParser()->getLoadManager()->calculate(managed.data(), int(managed.size())); }
					catch (...) { measurementError = std::current_exception(); }
				});

			int diverged = 0;
			const int rc = pffrg_compute_step(_handle, &diverged);
			if (measurementThread.joinable()) measurementThread.join();
			check(rc, "pffrg_compute_step");
			if (measurementError) std::rethrow_exception(measurementError);

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
			if (_options.syncState) downloadState();
			this->_flowingFunctional->cutoff = newCutoff;
		}

		// copy the device state into the reference's host arrays (needed before measurements / checkpoints when syncState is off)
		void downloadState()
		{
			void *v4[4] = { _state.v4[0], _state.v4[1], _state.v4[2], _state.v4[3] };
			double cutoff = 0.0;
			check(pffrg_get_state(_handle, &cutoff, _state.v2, v4, dtype()), "pffrg_get_state");
		}

		// correlations chi[c * L + rid] of `state` (the flowing functional), computed on the device; uploads the state first when
		// the host copy is the newer one (post-processing of deferred measurements reads checkpoints into the host arrays)
		void measureCorrelation(const EffectiveAction &state, std::vector<double> &chi)
		{
			if (&state != this->_flowingFunctional) throw Exception(Exception::Type::ArgumentError, "B200FrgCore::measureCorrelation: only the flowing functional can be measured");
			if (!_deviceCurrent || !((double)state.cutoff == _deviceCutoff)) uploadState();
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
			_deviceCurrent = true;
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
	};
}
