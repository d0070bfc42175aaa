// Round trip through spinparser_b200/host/hdf5_min.hpp alone (no reference code): a group with 300 children (two levels of B-tree nodes),
// an empty group, array-typed and plain datasets, attributes; written on H5Fclose, re-read by a fresh H5Fopen, extended and written again.
// Prints what tests/test_hdf5_min.py checks against the independent Python reader.
#include "hdf5_min.hpp"
#include <cstdio>

int main(int argc, char **argv)
{
	if (argc < 2) return 2;
	const char *name = argv[1];
	hid_t f = H5Fcreate(name, H5F_ACC_TRUNC, H5P_DEFAULT, H5P_DEFAULT);
	hid_t g = H5Gcreate(f, "Cor", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
	hid_t meta = H5Gcreate(g, "meta", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
	hsize_t three[1] = { 3 };
	hid_t vec = H5Tarray_create(H5T_NATIVE_FLOAT, 1, three);
	hsize_t nb[1] = { 2 };
	hid_t sp = H5Screate_simple(1, nb, nullptr);
	hid_t ds = H5Dcreate(meta, "basis", vec, sp, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
	H5MIN_REAL basis[6] = { 0, 0, 0, 0.5f, 0.25f, 1.5f };
	H5Dwrite(ds, vec, H5S_ALL, H5S_ALL, H5P_DEFAULT, basis);
	H5Dclose(ds); H5Sclose(sp); H5Tclose(vec); H5Gclose(meta);
	hid_t data = H5Gcreate(g, "data", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
	for (int k = 0; k < 300; ++k)
	{
		char nm[64]; std::snprintf(nm, sizeof nm, "measurement_%d", k);
		hid_t m = H5Gcreate(data, nm, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
		hsize_t one[1] = { 1 };
		hid_t as = H5Screate_simple(1, one, nullptr);
		hid_t a = H5Acreate(m, "cutoff", H5T_NATIVE_FLOAT, as, H5P_DEFAULT, H5P_DEFAULT);
		H5MIN_REAL c = (H5MIN_REAL)50 / (k + 1);
		H5Awrite(a, H5T_NATIVE_FLOAT, &c); H5Aclose(a); H5Sclose(as);
		hsize_t d2[2] = { 2, 5 };
		hid_t s2 = H5Screate_simple(2, d2, nullptr);
		hid_t dd = H5Dcreate(m, "data", H5T_NATIVE_FLOAT, s2, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
		H5MIN_REAL v[10]; for (int i = 0; i < 10; ++i) v[i] = (H5MIN_REAL)(k + 0.125 * i);
		H5Dwrite(dd, H5T_NATIVE_FLOAT, H5S_ALL, H5S_ALL, H5P_DEFAULT, v); H5Dclose(dd); H5Sclose(s2); H5Gclose(m);
	}
	H5Gclose(data); H5Gclose(g);
	hid_t e = H5Gcreate(f, "empty", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT); H5Gclose(e);
	if (H5Fclose(f) != 0) return 3;

	h5min::files().clear(); // a fresh process would see the disk only
	if (H5Fis_hdf5(name) <= 0 || H5Fis_hdf5("/nonexistent/file") > 0) return 4;
	hid_t r = H5Fopen(name, H5F_ACC_RDWR, H5P_DEFAULT);
	if (r < 0) return 5;
	hid_t d = H5Dopen(r, "Cor/data/measurement_299/data", H5P_DEFAULT);
	H5MIN_REAL v[10]; H5Dread(d, H5T_NATIVE_FLOAT, H5S_ALL, H5S_ALL, H5P_DEFAULT, v); H5Dclose(d);
	hid_t mg = H5Gopen(r, "Cor/data/measurement_17", H5P_DEFAULT);
	hid_t a = H5Aopen(mg, "cutoff", H5P_DEFAULT); H5MIN_REAL c; H5Aread(a, H5T_NATIVE_FLOAT, &c); H5Aclose(a); H5Gclose(mg);
	hsize_t n = 0; hid_t dg = H5Gopen(r, "Cor/data", H5P_DEFAULT); H5Gget_num_objs(dg, &n);
	char first[64]; H5Gget_objname_by_idx(dg, 0, first, sizeof first);
	const int type0 = H5Gget_objtype_by_idx(dg, 0);
	H5Gclose(dg);
	std::printf("%g %g %g %llu %s %d %d %d\n", (double)v[0], (double)v[9], (double)c, n, first, type0, (int)H5Lexists(r, "empty", H5P_DEFAULT), (int)H5Lexists(r, "Cor/nothing", H5P_DEFAULT));
	// extend the re-read file and store it again
	hid_t x = H5Gcreate(r, "extra", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT); H5Gclose(x);
	return H5Fclose(r) == 0 ? 0 : 6;
}
