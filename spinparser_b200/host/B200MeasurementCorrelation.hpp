// B200MeasurementCorrelation.hpp -- the `correlation` measurement of SpinParser with the susceptibility integral computed on
// the GPU (pffrg_measure_correlation) instead of the single-threaded work item of
// {SU2,XYZ,TRI}MeasurementCorrelation::_calculateCorrelation (src/SU2/SU2MeasurementCorrelation.cpp:77-177 and equivalents).
//
// Compiled inside the SpinParser tree next to B200FrgCore.hpp. The class derives from the reference's
// Measurement (src/Measurement.hpp:36-127), is created by FrgCoreFactory_b200.cpp for backend "b200", and writes the same
// .obs structure the reference writes (src/SU2/SU2MeasurementCorrelation.cpp:179-353):
//   /<Core>Cor<mu nu>/meta/{latticeVectors[3], basis[nb], sites[nb][nr]}  (float[3] each)
//   /<Core>Cor<mu nu>/data/measurement_<k>/{@cutoff, data[nb][nr]}
// through the HDF5 C API, so opt/python/spinparser/obs.py reads the files unchanged. The measurement is not load-managed:
// it needs neither the host copy of the vertex nor host threads.
#pragma once

#include <string>
#include <vector>

#include <hdf5.h>

#include "B200FrgCore.hpp"
#include "Measurement.hpp"

namespace b200
{
	// dataset names and the channel of chi[c][rid] each one reads, given the transformed spin components (sx, sy, sz)
	template <class RefCore> struct CorrelationLayout;
	template <> struct CorrelationLayout<SU2FrgCore>
	{
		static std::vector<std::string> names() { return { "SU2CorZZ", "SU2CorDD" }; }
		static int channel(int dataset, const int (&)[3]) { return dataset; }
	};
	template <> struct CorrelationLayout<XYZFrgCore>
	{
		static std::vector<std::string> names() { return { "XYZCorXX", "XYZCorYY", "XYZCorZZ", "XYZCorDD" }; }
		static int channel(int dataset, const int (&s)[3]) { return dataset < 3 ? s[dataset] : 3; }
	};
	template <> struct CorrelationLayout<TRIFrgCore>
	{
		static std::vector<std::string> names() { return { "TRICorXX", "TRICorXY", "TRICorXZ", "TRICorYX", "TRICorYY", "TRICorYZ", "TRICorZX", "TRICorZY", "TRICorZZ", "TRICorDD" }; }
		static int channel(int dataset, const int (&s)[3]) { return dataset < 9 ? 4 * s[dataset / 3] + s[dataset % 3] : 15; }
	};

	template <class RefCore>
	class B200MeasurementCorrelation : public Measurement, public DeviceMeasurement
	{
	public:
		B200MeasurementCorrelation(const std::string &outfile, const float minCutoff, const float maxCutoff, const bool defer)
			: Measurement(outfile, minCutoff, maxCutoff, defer, false)
		{
			const Lattice &l = FrgCommon::lattice();
			_nBasis = int(l._basis.size());
			_nRange = 0;
			for (auto i = l.getRange(0); i != l.end(); ++i) ++_nRange;
			// (basis site b, k-th site in range of b) -> representative and transformed spin components,
			// as src/TRI/TRIMeasurementCorrelation.cpp:297-307 evaluates them for every measurement
			for (auto i = l.getBasis(); i != l.end(); ++i)
				for (auto j = l.getRange(i); j != l.end(); ++j)
				{
					SpinComponent sx(SpinComponent::X), sy(SpinComponent::Y), sz(SpinComponent::Z);
					_rid.push_back(l.symmetryTransform(i, j, sx, sy, sz));
					_perm.push_back(static_cast<int>(sx)); _perm.push_back(static_cast<int>(sy)); _perm.push_back(static_cast<int>(sz));
				}
		}

		void takeMeasurement(const EffectiveAction &state, const bool isMasterTask) const override
		{
			B200FrgCore<RefCore> *core = dynamic_cast<B200FrgCore<RefCore> *>(SpinParser::spinParser()->getFrgCore());
			if (!core) throw Exception(Exception::Type::InternalError, "B200MeasurementCorrelation needs a B200FrgCore");
			const int L = FrgCommon::lattice().size;
			std::vector<double> chi;
			core->measureCorrelation(state, chi);
			if (!isMasterTask) return;
			const std::vector<std::string> names = CorrelationLayout<RefCore>::names();
			std::vector<real> buffer(_rid.size());
			for (size_t d = 0; d < names.size(); ++d)
			{
				for (size_t k = 0; k < _rid.size(); ++k)
				{
					const int s[3] = { _perm[3 * k], _perm[3 * k + 1], _perm[3 * k + 2] };
					buffer[k] = real(chi[size_t(CorrelationLayout<RefCore>::channel(int(d), s)) * L + _rid[k]]);
				}
				writeDataset(names[d], state.cutoff, buffer.data());
			}
		}

	private:
		void check(bool ok, const std::string &what) const { if (!ok) throw Exception(Exception::Type::IOError, what + " [" + outfile() + "]"); }

		// geometry of the measurement: lattice vectors, basis sites, and the sites the columns of every dataset refer to
		void writeMeta(hid_t group) const
		{
			const Lattice &l = FrgCommon::lattice();
			hid_t meta = H5Gcreate(group, "meta", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
			const hsize_t three[1] = { 3 };
			hid_t vec3 = H5Tarray_create(H5T_NATIVE_FLOAT, 1, three);
			auto put = [&](const char *name, int rank, const hsize_t *dims, const std::vector<real> &values)
			{
				hid_t space = H5Screate_simple(rank, dims, NULL);
				hid_t set = H5Dcreate(meta, name, vec3, space, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
				H5Dwrite(set, vec3, H5S_ALL, H5S_ALL, H5P_DEFAULT, values.data());
				H5Dclose(set); H5Sclose(space);
			};
			std::vector<real> v;
			for (auto &a : l._bravaisLattice) { v.push_back(real(a.x)); v.push_back(real(a.y)); v.push_back(real(a.z)); }
			const hsize_t nVectors[1] = { l._bravaisLattice.size() };
			put("latticeVectors", 1, nVectors, v);
			v.clear();
			for (auto b = l.getBasis(); b != l.end(); ++b) { auto p = l.getSitePosition(b); v.push_back(real(p.x)); v.push_back(real(p.y)); v.push_back(real(p.z)); }
			const hsize_t nBasis[1] = { hsize_t(_nBasis) };
			put("basis", 1, nBasis, v);
			v.clear();
			for (int b = 0; b < _nBasis; ++b)
				for (auto i = l.getRange(b); i != l.end(); ++i) { auto p = l.getSitePosition(i); v.push_back(real(p.x)); v.push_back(real(p.y)); v.push_back(real(p.z)); }
			const hsize_t nSites[2] = { hsize_t(_nBasis), hsize_t(_nRange) };
			put("sites", 2, nSites, v);
			H5Tclose(vec3);
			H5Gclose(meta);
		}

		void writeDataset(const std::string &observable, const real cutoff, const real *values) const
		{
			H5Eset_auto(H5E_DEFAULT, NULL, NULL);
			const std::string path = outfile();
			hid_t file = H5Fis_hdf5(path.c_str()) > 0 ? H5Fopen(path.c_str(), H5F_ACC_RDWR, H5P_DEFAULT) : H5Fcreate(path.c_str(), H5F_ACC_TRUNC, H5P_DEFAULT, H5P_DEFAULT);
			check(file >= 0, "Could not open observable file");
			hid_t group = H5Lexists(file, observable.c_str(), H5P_DEFAULT) > 0 ? H5Gopen(file, observable.c_str(), H5P_DEFAULT) : H5Gcreate(file, observable.c_str(), H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
			check(group >= 0, "Could not open obsfile group " + observable);
			if (H5Lexists(group, "meta", H5P_DEFAULT) == 0) writeMeta(group);
			hid_t data = H5Lexists(group, "data", H5P_DEFAULT) > 0 ? H5Gopen(group, "data", H5P_DEFAULT) : H5Gcreate(group, "data", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
			check(data >= 0, "Could not open obsfile group " + observable + "/data");

			// next free measurement id; a measurement at this cutoff that is already in the file is kept (duplicates are discarded)
			int id = 0;
			bool duplicate = false;
			hsize_t count = 0;
			H5Gget_num_objs(data, &count);
			for (hsize_t i = 0; i < count && !duplicate; ++i)
			{
				if (H5Gget_objtype_by_idx(data, i) != H5G_GROUP) continue;
				++id;
				char name[32];
				H5Gget_objname_by_idx(data, i, name, sizeof(name));
				hid_t existing = H5Gopen(data, name, H5P_DEFAULT);
				hid_t attribute = H5Aopen(existing, "cutoff", H5P_DEFAULT);
				real c = 0;
				H5Aread(attribute, H5T_NATIVE_FLOAT, &c);
				H5Aclose(attribute); H5Gclose(existing);
				duplicate = (c == cutoff);
			}
			if (duplicate) Log::log << Log::LogLevel::Warning << "Found existing correlation measurement at cutoff " + std::to_string(cutoff) + ". Discarding duplicate entry." << Log::endl;
			else
			{
				hid_t measurement = H5Gcreate(data, ("measurement_" + std::to_string(id)).c_str(), H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
				const hsize_t one[1] = { 1 };
				hid_t scalar = H5Screate_simple(1, one, NULL);
				hid_t attribute = H5Acreate(measurement, "cutoff", H5T_NATIVE_FLOAT, scalar, H5P_DEFAULT, H5P_DEFAULT);
				H5Awrite(attribute, H5T_NATIVE_FLOAT, &cutoff);
				H5Aclose(attribute); H5Sclose(scalar);
				const hsize_t dims[2] = { hsize_t(_nBasis), hsize_t(_nRange) };
				hid_t space = H5Screate_simple(2, dims, NULL);
				hid_t set = H5Dcreate(measurement, "data", H5T_NATIVE_FLOAT, space, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
				H5Dwrite(set, H5T_NATIVE_FLOAT, H5S_ALL, H5S_ALL, H5P_DEFAULT, values);
				H5Dclose(set); H5Sclose(space); H5Gclose(measurement);
			}
			H5Gclose(data); H5Gclose(group); H5Fclose(file);
		}

		int _nBasis, _nRange;
		std::vector<int> _rid, _perm;
	};
}
