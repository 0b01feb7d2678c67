// gpat_driver.cpp -- C++ host driver above the C ABI of libgpat_cuda.so.
//
// The reference's host side for this path is compiled Fortran (`program stochastic`,
// src/programs/stochastic-mhd.f90) and no Fortran compiler exists in this build image, so this
// is the compiled stand-in: the same command-line switches (stochastic-mhd.f90:576-1108, the
// subset the particle path uses; the others are accepted and must keep their defaults), the same
// conf.dat grammar (read_config.f90:22-45), the same mhd_config.dat / mhd_data_NNNN files
// (mhd_config.f90:139-148, mhd_data_parallel.f90:224-267) and the call sequence of
// solve_transport_equation (stochastic-mhd.f90:312-567) for one rank per GPU.  Everything
// numerical happens behind include/gpat_cuda.h.
//
// Outputs, in --diagnostics_directory: quick.dat and pmax_global.dat in the reference's text
// formats (diagnostics.f90:158-168, 1709-1718); the distributions as raw little-endian files
// (the reference's HDF5 writers, diagnostics.f90:1285-1642, stay Fortran -- no HDF5 here):
//   fdists_NNNN.bin       i32 nmu, npp | f64 fglobal(nmu,npp) | f64 pedges(npp+1) | f64 muedges(nmu+1)
//   fdists_localK_NNNN.bin  i32 nmu, np, nrx, nry, nrz | f64 flocalK(nmu,np,nrx,nry,nrz)
//
// build: make -C host      run: host/gpat_driver -dm <mhd dir>/ -cf conf.dat -np 100000 -te 3 ...
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../include/gpat_cuda.h"

namespace {

// ---- conf.dat: get_variable (read_config.f90:22-45) --------------------------------------------
struct ConfReader {
    std::vector<std::string> lines;
    size_t pos = 0;
    bool open(const std::string& path)
    {
        std::ifstream f(path);
        if (!f) return false;
        std::string l;
        while (std::getline(f, l)) lines.push_back(l);
        pos = 0;
        return true;
    }
    // forward-only scan for the first line containing `name`; value after '='; -1.0 if absent
    double get(const std::string& name)
    {
        while (pos < lines.size()) {
            const std::string& l = lines[pos++];
            if (l.find(name) == std::string::npos) continue;
            size_t eq = l.find('=');
            if (eq == std::string::npos) return -1.0;
            std::string v = l.substr(eq + 1);
            for (char& c : v)
                if (c == 'D' || c == 'd') c = 'E';
            return std::strtod(v.c_str(), nullptr);
        }
        return -1.0;
    }
};

// ---- command line (FLAP switches of stochastic-mhd.f90) ------------------------------------------
struct Cli {
    std::map<std::string, std::string> v;  // keyed by the short switch
    std::map<std::string, std::string> long2short;
    void def(const char* lng, const char* sht, const char* d) { v[sht] = d; long2short[lng] = sht; }
    bool parse(int argc, char** argv, std::string& err)
    {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            if (long2short.count(a)) a = long2short[a];
            if (!v.count(a)) { err = "unknown switch " + a; return false; }
            if (i + 1 >= argc) { err = "missing value for " + a; return false; }
            v[a] = argv[++i];
        }
        return true;
    }
    std::string s(const char* k) const { return v.at(k); }
    double d(const char* k) const { return std::strtod(v.at(k).c_str(), nullptr); }
    long long i(const char* k) const { return (long long)std::llround(std::strtod(v.at(k).c_str(), nullptr)); }
    bool b(const char* k) const
    {
        const std::string& x = v.at(k);
        return x == ".true." || x == ".TRUE." || x == "T" || x == "true" || x == "1";
    }
};

void define_switches(Cli& c)
{
    // defaults are the reference's (stochastic-mhd.f90:576-945)
    c.def("--quota_hour", "-qh", "5.00");            c.def("--restart_flag", "-rf", ".false.");
    c.def("--focused_transport", "-ft", ".false.");  c.def("--nlgc", "-nl", ".true.");
    c.def("--kperp_kpara", "-kk", "0.01");           c.def("--particle_v0", "-pv", "1.0");
    c.def("--size_mpi_sub", "-sm", "1");             c.def("--dir_mhd_data", "-dm", "");
    c.def("--mhd_config_filename", "-mc", "mhd_config.dat");
    c.def("--nptl_max", "-nm", "1E7");               c.def("--nptl", "-np", "1E4");
    c.def("--time_interp_flag", "-ti", "0");         c.def("--tinterval", "-dt", "1E-7");
    c.def("--tstart", "-ts", "0");                   c.def("--tend", "-te", "200");
    c.def("--tmax_mhd", "-tm", "100000");            c.def("--single_time_frame", "-st", "0");
    c.def("--dist_flag", "-df", "0");                c.def("--power_index", "-pi", "7.0");
    c.def("--split_flag", "-sf", "1");               c.def("--split_ratio", "-sr", "2.72");
    c.def("--pmin_split", "-ps", "2.0");             c.def("--track_particle_flag", "-tf", ".false.");
    c.def("--particle_tags_file", "-ptf", "");       c.def("--nsteps_interval", "-ni", "10");
    c.def("--diagnostics_directory", "-dd", "data/"); c.def("--inject_at_shock", "-is", ".false.");
    c.def("--inject_new_ptl", "-in", ".true.");      c.def("--inject_large_jz", "-ij", ".false.");
    c.def("--inject_same_nptl", "-sn", ".true.");    c.def("--tmax_to_inject", "-tti", "100000");
    c.def("--inject_part_box", "-ip", ".false.");    c.def("--jz_min", "-jz", "100.0");
    c.def("--ncells_large_jz_norm", "-nn", "800");   c.def("--inject_large_db2", "-ib", ".false.");
    c.def("--db2_min", "-db2", "0.03");              c.def("--ncells_large_db2_norm", "-nb", "800");
    c.def("--inject_large_divv", "-iv", ".false.");  c.def("--divv_min", "-dv", "10.0");
    c.def("--ncells_large_divv_norm", "-nv", "800"); c.def("--inject_large_rho", "-ir", ".false.");
    c.def("--rho_min", "-rm", "2.0");                c.def("--ncells_large_rho_norm", "-nr", "800");
    c.def("--inject_large_absj", "-iaj", ".false."); c.def("--absj_min", "-ajm", "100.0");
    c.def("--ncells_large_absj_norm", "-naj", "800");
    c.def("--ptl_xmin", "-xs", "0.0");               c.def("--ptl_xmax", "-xe", "1.0");
    c.def("--ptl_ymin", "-ys", "0.0");               c.def("--ptl_ymax", "-ye", "1.0");
    c.def("--ptl_zmin", "-zs", "0.0");               c.def("--ptl_zmax", "-ze", "1.0");
    c.def("--conf_file", "-cf", "conf.dat");         c.def("--num_fine_steps", "-nf", "1");
    c.def("--local_dist", "-ld", ".true.");          c.def("--dpp_wave", "-dw", "0");
    c.def("--dpp_shear", "-ds", "0");                c.def("--weak_scattering", "-ws", "1");
    c.def("--tau0_scattering", "-t0", "1.0");        c.def("--deltab_flag", "-db", "0");
    c.def("--correlation", "-co", "0");              c.def("--ndim_field", "-nd", "2");
    c.def("--drift_param1", "-dp1", "4E7");          c.def("--drift_param2", "-dp2", "2E8");
    c.def("--charge", "-ch", "-1");                  c.def("--spherical_coord", "-sc", "0");
    c.def("--uniform_grid", "-ug", "1");             c.def("--check_drift_2d", "-cd", "0");
    c.def("--particle_data_dump", "-pd", "0");       c.def("--include_3rd_dim", "-i3", "0");
    c.def("--acc_by_surface", "-as", "0");           c.def("--surface_filename1", "-sf1", "");
    c.def("--surface_norm1", "-sn1", "+y");          c.def("--surface2_existed", "-s2e", ".false.");
    c.def("--is_intersection", "-ii", ".false.");    c.def("--surface_filename2", "-sf2", "");
    c.def("--surface_norm2", "-sn2", "-y");          c.def("--varying_dt_mhd", "-vdt", ".false.");
    c.def("--duu_init", "-du", "1.0");               c.def("--dump_escaped_dist", "-ded", ".false.");
    c.def("--dump_escaped", "-de", ".false.");
    // this driver only
    c.def("--device", "-gpu", "0");                  c.def("--seed", "-seed", "98443300");  // 0x5DE2024
    c.def("--strict_math", "-strict", "0");
}

// ---- mhd_config.dat: 13 f64 + 13 i32 (mhd_config.f90:139-148) --------------------------------------
struct MhdConfig {
    double dx, dy, dz, xmin, ymin, zmin, xmax, ymax, zmax, lx, ly, lz, dt_out;
    int32_t nx, ny, nz, nxs, nys, nzs, topox, topoy, topoz, nvar, bcx, bcy, bcz;
};

bool read_mhd_config(const std::string& path, MhdConfig& m)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    bool ok = std::fread(&m.dx, sizeof(double), 13, f) == 13 && std::fread(&m.nx, sizeof(int32_t), 13, f) == 13;
    std::fclose(f);
    return ok;
}

bool read_frame(const std::string& dir, int frame, size_t nfloats, std::vector<float>& buf,
                const char* stem = "mhd_data")
{
    char name[48];
    std::snprintf(name, sizeof(name), "%s_%04d", stem, frame);
    FILE* f = std::fopen((dir + name).c_str(), "rb");
    if (!f) return false;
    buf.resize(nfloats);
    bool ok = std::fread(buf.data(), sizeof(float), nfloats, f) == nfloats;
    std::fclose(f);
    return ok;
}

// Fortran edit descriptor E13.6: 0.dddddd E+xx, right-justified in 13 columns
std::string fortran_e13_6(double v)
{
    char out[64];
    if (v == 0.0) return " 0.000000E+00";
    int e = (int)std::floor(std::log10(std::fabs(v))) + 1;
    double m = v / std::pow(10.0, e);
    if (std::fabs(m) >= 0.9999995) { m /= 10.0; e += 1; }
    std::snprintf(out, sizeof(out), "%s0.%06lldE%+03d", v < 0 ? "-" : " ",
                  (long long)std::llround(std::fabs(m) * 1e6), e);
    return out;
}

int die(gpat_handle h, const char* what, int rc)
{
    // the reference prints and stops (simulation_setup.f90:78-87)
    std::fprintf(stderr, "gpat_driver: %s failed (%d): %s\n", what, rc, gpat_last_error(h));
    if (h) gpat_finalize(h);
    return 1;
}

}  // namespace

#define CK(call, what)                    \
    do {                                  \
        int rc_ = (call);                 \
        if (rc_) return die(h, what, rc_); \
    } while (0)

int main(int argc, char** argv)
{
    Cli cli;
    define_switches(cli);
    std::string err;
    if (!cli.parse(argc, argv, err)) {
        std::fprintf(stderr, "gpat_driver: %s\n", err.c_str());
        return 2;
    }
    const std::string dir_mhd = cli.s("-dm"), diag_dir = cli.s("-dd");
    MhdConfig mc{};
    if (!read_mhd_config(dir_mhd + cli.s("-mc"), mc)) {
        std::fprintf(stderr, "gpat_driver: cannot read %s%s\n", dir_mhd.c_str(), cli.s("-mc").c_str());
        return 2;
    }
    // features outside the GPU path must stay off; the library re-checks the ones it is told about
    for (const char* k : {"-sc"})
        if (cli.b(k)) {
            std::fprintf(stderr, "gpat_driver: switch %s is outside the GPU particle path\n", k);
            return 2;
        }

    gpat_params P{};
    P.ndim = (int)cli.i("-nd");
    P.nx = mc.nx; P.ny = mc.ny; P.nz = mc.nz;
    P.time_interp = (int)cli.i("-ti");
    P.dx = mc.dx; P.dy = mc.dy; P.dz = mc.dz;
    P.xmin = mc.xmin; P.ymin = mc.ymin; P.zmin = mc.zmin;  // fconfig for a 1x1x1 topology, simulation_setup.f90:171-253
    P.xmax = mc.xmax; P.ymax = mc.ymax; P.zmax = mc.zmax;
    P.lx = mc.lx; P.ly = mc.ly; P.lz = mc.lz;
    const int t_start = (int)cli.i("-ts"), t_end = (int)cli.i("-te");
    const int nframes_run = t_end - t_start;

    // read_particle_params: one open, keys in this order (particle_module.f90:2790-2814)
    ConfReader r;
    if (!r.open(cli.s("-cf"))) {
        std::fprintf(stderr, "gpat_driver: cannot read %s\n", cli.s("-cf").c_str());
        return 2;
    }
    P.b0 = r.get("b0"); P.p0 = r.get("p0"); P.pmin = r.get("pmin"); P.pmax = r.get("pmax");
    P.momentum_dependency = (int)r.get("momentum_dependency");
    P.gamma_turb = r.get("gamma_turb");
    P.pindex = 3.0 - P.gamma_turb;
    P.mag_dependency = (int)r.get("mag_dependency");
    P.kpara0 = r.get("kpara0"); P.kret = r.get("kret");
    P.dt_min_rel = r.get("dt_min_rel"); P.dt_max_rel = r.get("dt_max_rel");
    P.acc_region_flag = (int)r.get("acc_region_flag");
    const char* acc_keys[6] = {"acc_xmin", "acc_xmax", "acc_ymin", "acc_ymax", "acc_zmin", "acc_zmax"};
    for (int k = 0; k < 6; ++k) P.acc_region[k] = r.get(acc_keys[k]);
    if (P.acc_region_flag != 1) {  // particle_module.f90:2868-2877
        const double whole[6] = {0.0, 1.0, 0.0, 1.0, 0.0, 1.0};
        std::memcpy(P.acc_region, whole, sizeof(whole));
    }
    // read_diagnostics_params: a fresh open (diagnostics.f90:2060-2104)
    r.pos = 0;
    P.npp_global = (int)r.get("npp_global");
    const bool ft = cli.b("-ft");
    const int nmu_global_conf = (int)r.get("nmu_global");
    P.nmu_global = ft ? nmu_global_conf : 1;  // 1 for Parker transport, diagnostics.f90:2107-2111
    for (int k = 0; k < 4; ++k) {
        const std::string n = std::to_string(k + 1);
        gpat_hist_spec& s = P.local[k];
        const int dump_interval = (int)r.get("dump_interval" + n);
        s.pmin = r.get("pmin" + n); s.pmax = r.get("pmax" + n);
        s.npbins = (int)r.get("npbins" + n);
        const int nmu_conf = (int)r.get("nmu" + n);
        s.nmu = ft ? nmu_conf : 1;  // diagnostics.f90:2124-2128
        s.rx = (int)r.get("rx" + n); s.ry = (int)r.get("ry" + n); s.rz = (int)r.get("rz" + n);
        s.enabled = dump_interval < nframes_run ? 1 : 0;  // nframes = t_end - t_start, diagnostics.f90:2129, stochastic-mhd.f90:205
    }
    // read_particle_boundary_conditions (simulation_setup.f90:107-123)
    r.pos = 0;
    P.pbc[0] = (int)r.get("pbcx"); P.pbc[1] = (int)r.get("pbcy"); P.pbc[2] = (int)r.get("pbcz");

    P.dpp_wave = (int)cli.i("-dw"); P.dpp_shear = (int)cli.i("-ds"); P.weak_scattering = (int)cli.i("-ws");
    P.tau0 = cli.d("-t0");
    P.drift1 = cli.d("-dp1"); P.drift2 = cli.d("-dp2"); P.pcharge = (int)cli.i("-ch");
    P.check_drift_2d = (int)cli.i("-cd"); P.include_3rd_dim = (int)cli.i("-i3");
    P.nlgc = cli.b("-nl") ? 1 : 0; P.kperp_kpara = cli.d("-kk");
    P.focused_transport = ft ? 1 : 0;
    P.duu0 = cli.d("-du");  // set_duu_params, stochastic-mhd.f90:207
    P.spherical_coord = (int)cli.i("-sc"); P.nonuniform_grid = 1 - (int)cli.i("-ug");
    P.deltab_flag = (int)cli.i("-db"); P.correlation_flag = (int)cli.i("-co"); P.acc_by_surface = (int)cli.i("-as");
    // surface_norm1/2 ('+x' ... '-z', stochastic-mhd.f90:911-938): anything but x / y in the second character
    // is z and anything but '+' in the first is the negative direction (acc_region_surface.f90:44-50, 350-366)
    auto norm_code = [](const std::string& v) {
        const char sign = v.size() > 0 ? v[0] : ' ', ax = v.size() > 1 ? v[1] : ' ';
        const int axis = ax == 'x' ? 1 : (ax == 'y' ? 2 : 3);
        return sign == '+' ? axis : -axis;
    };
    P.surface_norm1 = norm_code(cli.s("-sn1")); P.surface_norm2 = norm_code(cli.s("-sn2"));
    P.surface2_existed = cli.b("-s2e") ? 1 : 0; P.is_intersection = cli.b("-ii") ? 1 : 0;
    P.seed = (uint64_t)cli.i("-seed"); P.rng_mode = GPAT_RNG_PHILOX; P.mpi_rank = 0;
    P.strict_math = (int)cli.i("-strict");
    P.keep_rho = cli.b("-ir") ? 1 : 0;  // inject_large_rho interpolates the density (particle_module.f90:1448)

    gpat_handle h = nullptr;
    const long long nptl_max = cli.i("-nm"), nptl = cli.i("-np");
    CK(gpat_init(&h, (int)cli.i("-gpu"), nptl_max, &P), "gpat_init");

    // restart_flag (stochastic-mhd.f90:224-239): tmin from restart/latest_restart, then read_particles(tmin)
    // and read_particle_module_state(tmin).  The reference's HDF5 containers are raw records here (the
    // formats are stated next to stochastic_parker_b200.driver.dump_restart, which writes the same files).
    int tmin = t_start;
    if (cli.b("-rf")) {
        const std::string rdir = diag_dir + "restart/";
        int32_t t32 = 0;
        FILE* f = std::fopen((rdir + "latest_restart").c_str(), "rb");
        if (!f || std::fread(&t32, sizeof(t32), 1, f) != 1) return die(h, "read restart/latest_restart", -1);
        std::fclose(f);
        tmin = t32;
        char name[64];
        std::snprintf(name, sizeof(name), "particles_%04d.bin", tmin);
        f = std::fopen((rdir + name).c_str(), "rb");
        int64_t n = 0;
        if (!f || std::fread(&n, sizeof(n), 1, f) != 1 || n < 0 || n > nptl_max) return die(h, "read restart particles header", -1);
        std::vector<gpat_particle> ptl((size_t)n);
        if (std::fread(ptl.data(), sizeof(gpat_particle), (size_t)n, f) != (size_t)n) return die(h, "read restart particles", -1);
        std::fclose(f);
        std::snprintf(name, sizeof(name), "particle_module_state_%04d.bin", tmin);
        gpat_counters rc{};
        f = std::fopen((rdir + name).c_str(), "rb");
        if (!f || std::fread(&rc, sizeof(rc), 1, f) != 1) return die(h, "read restart particle_module_state", -1);
        std::fclose(f);
        CK(gpat_upload_particles(h, ptl.data(), n), "gpat_upload_particles");
        CK(gpat_set_counters(h, &rc), "gpat_set_counters");
        std::printf("This is a restart of a previous simulation (frame %d, %lld particles)\n", tmin, (long long)n);
    }

    // farray(:, -1:nx+2, [-1:ny+2, [-1:nz+2]]) (mhd_data_parallel.f90:77-83)
    const size_t ncell = (size_t)(mc.nx + 4) * (P.ndim >= 2 ? mc.ny + 4 : 1) * (P.ndim == 3 ? mc.nz + 4 : 1);
    std::vector<float> frame;
    // calc_tstamps_mhd (mhd_config.f90:263-271): uniform output interval
    // or, with -vdt .true., load_tstamps_mhd (mhd_config.f90:221-254): time_stamps.dat = the first and last frame of
    // the MHD run in two 8-byte slots (the first is read as a default integer), then one f64 per frame; frames past
    // tmax_mhd continue with the last interval
    std::vector<double> stamps((size_t)(t_end - t_start + 1));
    for (int i = t_start; i <= t_end; ++i) stamps[(size_t)(i - t_start)] = i * mc.dt_out;
    if (cli.b("-vdt")) {
        FILE* f = std::fopen((dir_mhd + "time_stamps.dat").c_str(), "rb");
        int32_t ts_mhd = 0;
        if (!f || std::fread(&ts_mhd, sizeof(ts_mhd), 1, f) != 1) return die(h, "read time_stamps.dat", -1);
        const long long tm = cli.i("-tm");
        const long long nread = std::min<long long>(tm, t_end) - t_start + 1;
        if (nread < 2 || std::fseek(f, (long)(t_start - ts_mhd + 2) * 8, SEEK_SET) != 0 ||
            std::fread(stamps.data(), sizeof(double), (size_t)nread, f) != (size_t)nread)
            return die(h, "read time_stamps.dat", -1);
        std::fclose(f);
        for (long long i = nread; i <= t_end - t_start; ++i) stamps[(size_t)i] = stamps[(size_t)i - 1] + (stamps[(size_t)nread - 1] - stamps[(size_t)nread - 2]);
    }
    auto tstamp = [&](int i) { return stamps[(size_t)(i - t_start)]; };
    double part_box[6] = {P.xmin, P.ymin, P.zmin, P.xmax, P.ymax, P.zmax};  // stochastic-mhd.f90:375-391
    if (cli.b("-ip")) {
        part_box[0] = cli.d("-xs"); part_box[1] = cli.d("-ys"); part_box[2] = cli.d("-zs");
        part_box[3] = cli.d("-xe"); part_box[4] = cli.d("-ye"); part_box[5] = cli.d("-ze");
    }
    const int dist_flag = (int)cli.i("-df"), split_flag = (int)cli.i("-sf"), nsteps_interval = (int)cli.i("-ni");
    const int num_fine_steps = (int)cli.i("-nf"), single_frame = (int)cli.i("-st");
    const bool local_dist = cli.b("-ld"), dump_escaped_dist = cli.b("-ded"), inject_new = cli.b("-in");
    const long long tmax_to_inject = cli.i("-tti"), tmax_mhd = cli.i("-tm");

    // diagnostics buffers with the reference's shapes (diagnostics.f90:182-191, 235-245)
    std::vector<double> fglobal((size_t)P.nmu_global * P.npp_global), pedges(P.npp_global + 1), muedges(P.nmu_global + 1);
    std::vector<double> flocal[4];
    int lshape[4][5];
    double* flocal_ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; ++k) {
        const gpat_hist_spec& s = P.local[k];
        if (!s.enabled) continue;
        const int nrx = (P.nx + s.rx - 1) / s.rx, nry = (P.ny + s.ry - 1) / s.ry, nrz = (P.nz + s.rz - 1) / s.rz;
        const int shp[5] = {s.nmu, s.npbins, nrx, nry, nrz};
        std::memcpy(lshape[k], shp, sizeof(shp));
        flocal[k].assign((size_t)s.nmu * s.npbins * nrx * nry * nrz, 0.0);
        flocal_ptr[k] = flocal[k].data();
    }
    CK(gpat_hist_edges(h, 0, pedges.data(), muedges.data()), "gpat_hist_edges");

    auto diagnostics = [&](int iframe, bool create) -> int {
        double quick[8], pmax = 0.0;
        int rc = gpat_diagnostics(h, local_dist ? 1 : 0, fglobal.data(), flocal_ptr, quick, &pmax);
        if (rc) return rc;
        // quick_check, diagnostics.f90:158-168
        FILE* f = std::fopen((diag_dir + "quick.dat").c_str(), create ? "w" : "a");
        if (!f) return -1;
        if (create)
            std::fprintf(f, "%6s%13s%13s%13s%13s%13s%13s%13s%13s\n", "iframe", "nptl_current", "nptl_split", "ntot",
                         "leak", "leak_negp", "pdt_min", "pdt_max", "pdt_avg");
        const double avg = quick[0] > 0 ? quick[5] / quick[0] : 0.0;
        std::fprintf(f, "%06d", iframe);
        for (int k = 0; k < 5; ++k) std::fputs(fortran_e13_6(quick[k]).c_str(), f);
        std::fputs(fortran_e13_6(quick[6]).c_str(), f);
        std::fputs(fortran_e13_6(quick[7]).c_str(), f);
        std::fputs(fortran_e13_6(avg).c_str(), f);
        std::fputc('\n', f);
        std::fclose(f);
        // get_pmax_global, diagnostics.f90:1709-1718
        f = std::fopen((diag_dir + "pmax_global.dat").c_str(), create ? "w" : "a");
        if (!f) return -1;
        std::fprintf(f, "%s\n", fortran_e13_6(pmax).c_str());
        std::fclose(f);
        // distributions (raw stand-in of save_global/local_distributions)
        char name[64];
        std::snprintf(name, sizeof(name), "fdists_%04d.bin", iframe);
        f = std::fopen((diag_dir + name).c_str(), "wb");
        if (!f) return -1;
        const int32_t hdr[2] = {P.nmu_global, P.npp_global};
        std::fwrite(hdr, sizeof(int32_t), 2, f);
        std::fwrite(fglobal.data(), sizeof(double), fglobal.size(), f);
        std::fwrite(pedges.data(), sizeof(double), pedges.size(), f);
        std::fwrite(muedges.data(), sizeof(double), muedges.size(), f);
        std::fclose(f);
        for (int k = 0; k < 4 && local_dist; ++k) {
            if (!flocal_ptr[k]) continue;
            std::snprintf(name, sizeof(name), "fdists_local%d_%04d.bin", k + 1, iframe);
            f = std::fopen((diag_dir + name).c_str(), "wb");
            if (!f) return -1;
            std::fwrite(lshape[k], sizeof(int32_t), 5, f);
            std::fwrite(flocal[k].data(), sizeof(double), flocal[k].size(), f);
            std::fclose(f);
        }
        return 0;
    };

    // dump_particles (-pd 1, diagnostics.f90:1811-1888) and dump_escaped_particles (-de .true. with -ded,
    // :1896-1978): int64 count + gpat_particle records, the format of the restart files
    auto write_particles = [&](const std::string& path, bool escaped) -> int {
        gpat_counters cc{};
        int rc = gpat_get_counters(h, &cc);
        if (rc) return rc;
        const int64_t cap = escaped ? cc.nptl_escaped : cc.nptl_current;
        std::vector<gpat_particle> buf((size_t)std::max<int64_t>(cap, 1));
        int64_t n = 0;
        rc = escaped ? gpat_download_escaped(h, buf.data(), cap, &n) : gpat_download_particles(h, buf.data(), cap, &n);
        if (rc) return rc;
        n = std::min(n, cap);
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) return -1;
        std::fwrite(&n, sizeof(n), 1, f);
        std::fwrite(buf.data(), sizeof(gpat_particle), (size_t)n, f);
        std::fclose(f);
        return 0;
    };
    const bool particle_data_dump = cli.i("-pd") == 1, dump_escaped = cli.b("-de");
    auto frame_name = [](const char* stem, int iframe) {
        char name[64];
        std::snprintf(name, sizeof(name), "%s_%04d.bin", stem, iframe);
        return std::string(name);
    };

    // escaped_dists_NNNN: calc_escaped_distributions + save_global/local_escaped_distributions
    // (diagnostics.f90:913-1232, 1331-1372, 1536-1642) as raw records:
    //   escaped_dists_NNNN.bin         int32 nmu, npp, nface; fescaped(nmu, npp, nface)
    //   escaped_dists_localK_NNNN.bin  int32 nmu, npbins, nrx, nry, nrz, ndim; fescapedK_x, [_y, [_z]]
    auto escaped_diagnostics = [&](int iframe) -> int {
        const int nface = 2 * P.ndim;
        std::vector<double> fesc((size_t)P.nmu_global * P.npp_global * nface);
        int rc = gpat_escaped_diagnostics(h, fesc.data());
        if (rc) return rc;
        char name[64];
        std::snprintf(name, sizeof(name), "escaped_dists_%04d.bin", iframe);
        FILE* f = std::fopen((diag_dir + name).c_str(), "wb");
        if (!f) return -1;
        const int32_t hdr[3] = {P.nmu_global, P.npp_global, nface};
        std::fwrite(hdr, sizeof(int32_t), 3, f);
        std::fwrite(fesc.data(), sizeof(double), fesc.size(), f);
        std::fclose(f);
        if (!local_dist) return 0;
        std::vector<double> face[3][4];
        double *fx[4] = {}, *fy[4] = {}, *fz[4] = {};
        for (int k = 0; k < 4; ++k) {
            if (!flocal_ptr[k]) continue;
            const size_t nb = (size_t)lshape[k][0] * lshape[k][1] * 2;
            face[0][k].assign(nb * lshape[k][3] * lshape[k][4], 0.0); fx[k] = face[0][k].data();
            if (P.ndim > 1) { face[1][k].assign(nb * lshape[k][2] * lshape[k][4], 0.0); fy[k] = face[1][k].data(); }
            if (P.ndim > 2) { face[2][k].assign(nb * lshape[k][2] * lshape[k][3], 0.0); fz[k] = face[2][k].data(); }
        }
        rc = gpat_escaped_local_diagnostics(h, fx, fy, fz);
        if (rc) return rc;
        for (int k = 0; k < 4; ++k) {
            if (!flocal_ptr[k]) continue;
            std::snprintf(name, sizeof(name), "escaped_dists_local%d_%04d.bin", k + 1, iframe);
            f = std::fopen((diag_dir + name).c_str(), "wb");
            if (!f) return -1;
            std::fwrite(lshape[k], sizeof(int32_t), 5, f);
            const int32_t nd = P.ndim;
            std::fwrite(&nd, sizeof(int32_t), 1, f);
            for (int a = 0; a < 3; ++a) std::fwrite(face[a][k].data(), sizeof(double), face[a][k].size(), f);
            std::fclose(f);
        }
        return 0;
    };

    // ---- particle tracking (stochastic-mhd.f90:226-229): the tag table is the reference's HDF5
    // dataset "tags" (nptl_tracking x (split_times_max+2) int32, C order) as a raw file
    // [int32 nptl_tracking, int32 ncols, data...] -- no HDF5 in this image
    const bool track = cli.b("-tf");
    if (track) {
        FILE* f = std::fopen(cli.s("-ptf").c_str(), "rb");
        int32_t hdr[2] = {0, 0};
        if (!f || std::fread(hdr, sizeof(int32_t), 2, f) != 2 || hdr[0] < 1 || hdr[1] < 2)
            return die(h, "read particle_tags_file header", -1);
        std::vector<int32_t> tags((size_t)hdr[0] * hdr[1]);
        if (std::fread(tags.data(), sizeof(int32_t), tags.size(), f) != tags.size())
            return die(h, "read particle_tags_file", -1);
        std::fclose(f);
        CK(gpat_init_tracking(h, tags.data(), hdr[1], hdr[0], nsteps_interval), "gpat_init_tracking");
    }
    auto dump_tracked = [&](int iframe) -> int {  // dump_tracked_particles, particle_module.f90:6236-6299
        int64_t nmax = 0, ntrk = 0;
        int rc = gpat_tracked_shape(h, &nmax, &ntrk);
        if (rc) return rc;
        std::vector<gpat_particle> rec((size_t)nmax * ntrk);
        rc = gpat_download_tracked(h, rec.data());
        if (rc) return rc;
        char name[96];
        std::snprintf(name, sizeof(name), "particle_tracking_particles_tracked_%04d.bin", iframe);
        FILE* f = std::fopen((diag_dir + name).c_str(), "wb");
        if (!f) return -1;
        const int64_t hdr[2] = {ntrk, nmax};
        std::fwrite(hdr, sizeof(int64_t), 2, f);
        std::fwrite(rec.data(), sizeof(gpat_particle), rec.size(), f);
        std::fclose(f);
        return gpat_reset_tracked(h);
    };

    // a missing or short input file: say which, and let CK() finalize the handle once
    auto file_error = [](const char* what, int tframe) -> int {
        std::fprintf(stderr, "gpat_driver: cannot read %s (frame %d)\n", what, tframe);
        return -1;
    };

    // deltab_NNNN / lc_NNNN: slab array then 2-D array (read_magnetic_fluctuation,
    // read_correlation_length, mhd_data_parallel.f90:306-497; stochastic-mhd.f90:330-346, 413-420)
    std::vector<float> maps;
    auto upload_maps = [&](int tframe, int slot) -> int {
        if (P.deltab_flag) {
            if (!read_frame(dir_mhd, tframe, ncell * 2, maps, "deltab")) return file_error("deltab", tframe);
            int rc = gpat_upload_turbulence(h, 0, slot, maps.data());
            if (rc) return rc;
        }
        if (P.correlation_flag) {
            if (!read_frame(dir_mhd, tframe, ncell * 2, maps, "lc")) return file_error("lc", tframe);
            int rc = gpat_upload_turbulence(h, 1, slot, maps.data());
            if (rc) return rc;
        }
        return 0;
    };

    // <surface_filenameK>_NNNN.dat: float64 heights over the ghosted plane (read_acc_surface,
    // acc_region_surface.f90:118-206; stochastic-mhd.f90:323-334, 405-416)
    std::vector<double> surf;
    auto upload_surfaces = [&](int tframe, int slot) -> int {
        if (!P.acc_by_surface) return 0;
        for (int k = 0; k < (P.surface2_existed ? 2 : 1); ++k) {
            const int axis = std::abs(k ? P.surface_norm2 : P.surface_norm1) - 1;
            const size_t n = (size_t)(axis == 0 ? mc.ny + 4 : mc.nx + 4) * (axis == 2 ? mc.ny + 4 : mc.nz + 4);
            char name[512];
            std::snprintf(name, sizeof(name), "%s%s_%04d.dat", dir_mhd.c_str(), cli.s(k ? "-sf2" : "-sf1").c_str(), tframe);
            surf.resize(n);
            FILE* f = std::fopen(name, "rb");
            if (!f) return file_error(name, tframe);
            const size_t got = std::fread(surf.data(), sizeof(double), n, f);
            std::fclose(f);
            if (got != n) return file_error(name, tframe);
            int rc = gpat_upload_acc_surface(h, k, slot, surf.data());
            if (rc) return rc;
        }
        return 0;
    };

    // ---- solve_transport_equation (stochastic-mhd.f90:312-567) ----
    if (!read_frame(dir_mhd, tmin, ncell * 8, frame)) return die(h, "read first mhd_data frame", -1);
    CK(gpat_upload_fields(h, 0, frame.data(), 8, 0), "gpat_upload_fields");
    CK(upload_surfaces(tmin, 0), "gpat_upload_acc_surface");
    CK(upload_maps(tmin, 0), "gpat_upload_turbulence");
    uint64_t total_steps = 0;
    auto wall0 = std::chrono::steady_clock::now();
    auto step1 = wall0;
    bool reached_quota = false;
    int last_read = tmin;
    int tf = tmin + 1;
    for (; tf <= t_end; ++tf) {
        std::printf(" Starting step %d\n", tf);
        if (single_frame == 0 && tf <= tmax_mhd) {  // :400-447
            if (!read_frame(dir_mhd, tf, ncell * 8, frame)) return die(h, "read mhd_data frame", -1);
            CK(gpat_upload_fields(h, P.time_interp ? 1 : 0, frame.data(), 8, 0), "gpat_upload_fields");
            CK(upload_surfaces(tf, P.time_interp ? 1 : 0), "gpat_upload_acc_surface");
            CK(upload_maps(tf, P.time_interp ? 1 : 0), "gpat_upload_turbulence");
            last_read = tf;
        } else if (P.time_interp && tf > tmin + 1) {
            // no new frame (tf > tmax_mhd): the reference's farray2 still holds the last frame it read, and
            // copy_fields made farray1 equal to it.  gpat_swap_fields exchanges the two device halves instead of
            // copying, so the last frame is sent to slot 1 again.
            CK(gpat_upload_fields(h, 1, frame.data(), 8, 0), "gpat_upload_fields");
            CK(upload_surfaces(last_read, 1), "gpat_upload_acc_surface");
            CK(upload_maps(last_read, 1), "gpat_upload_turbulence");
        }
        const double t0 = tstamp(tf - 1), dtf = tstamp(tf) - tstamp(tf - 1);  // tstamps_mhd(tf - t_start), particle_module.f90:1868
        if (cli.b("-is")) {  // :451-454: locate_shock_xpos + inject_particles_at_shock, every frame
            CK(gpat_inject_at_shock(h, nptl, cli.d("-dt"), dist_flag, cli.d("-pv"), t0, cli.d("-pi")),
               "gpat_inject_at_shock");
        } else if ((tf == t_start + 1 || inject_new) && tf <= tmax_to_inject) {  // :462-485, same precedence
            int mode = 0;
            double vmin = 0.0;
            long long norm = 1;
            if (cli.b("-ij")) { mode = GPAT_INJECT_LARGE_JZ; vmin = cli.d("-jz"); norm = cli.i("-nn"); }
            else if (cli.b("-iaj")) { mode = GPAT_INJECT_LARGE_ABSJ; vmin = cli.d("-ajm"); norm = cli.i("-naj"); }
            else if (cli.b("-ib")) { mode = GPAT_INJECT_LARGE_DB2; vmin = cli.d("-db2"); norm = cli.i("-nb"); }
            else if (cli.b("-iv")) { mode = GPAT_INJECT_LARGE_DIVV; vmin = cli.d("-dv"); norm = cli.i("-nv"); }
            else if (cli.b("-ir")) { mode = GPAT_INJECT_LARGE_RHO; vmin = cli.d("-rm"); norm = cli.i("-nr"); }
            if (mode) {
                int64_t ninj = 0, ncells = 0;
                CK(gpat_inject_targeted(h, mode, nptl, cli.d("-dt"), dist_flag, cli.d("-pv"), t0, dtf, part_box,
                                        cli.d("-pi"), cli.b("-sn") ? 1 : 0, vmin, norm, &ninj, &ncells),
                   "gpat_inject_targeted");
            } else {
                CK(gpat_inject_uniform(h, nptl, cli.d("-dt"), dist_flag, cli.d("-pv"), t0, dtf, part_box, cli.d("-pi")),
                   "gpat_inject_uniform");
            }
        }
        if (tf == t_start + 1 && !track) {  // :488-494
            CK(diagnostics(t_start, true), "initial diagnostics");
            if (particle_data_dump) CK(write_particles(diag_dir + frame_name("particles", t_start), false), "dump_particles");
        }
        uint64_t steps = 0;
        CK(gpat_particle_mover(h, t0, dtf, nsteps_interval, track ? 1 : num_fine_steps, dump_escaped_dist ? 1 : 0,
                               &steps),
           "gpat_particle_mover");  // :497-503: a tracking run moves with num_fine_steps = 1
        total_steps += steps;
        std::printf(" Finishing moving particles \n");
        if (track) CK(dump_tracked(tf), "dump_tracked_particles");  // :509-511
        if (split_flag == 1) CK(gpat_split(h, cli.d("-sr"), cli.d("-ps"), nsteps_interval), "gpat_split");  // :515
        if (!track) {  // :516-535
            CK(diagnostics(tf, false), "diagnostics");  // :518-521
            std::printf(" Finishing distribution diagnostics \n");
            if (particle_data_dump) CK(write_particles(diag_dir + frame_name("particles", tf), false), "dump_particles");  // :525-527
            if (dump_escaped_dist) {  // :522-534: escaped spectra of this interval, then reset_escaped_particles
                CK(escaped_diagnostics(tf), "escaped diagnostics");
                if (dump_escaped) CK(write_particles(diag_dir + frame_name("escaped_particles", tf), true), "dump_escaped_particles");
                CK(gpat_reset_escaped(h), "gpat_reset_escaped");
            }
        }
        if (P.time_interp == 1) {
            CK(gpat_swap_fields(h), "gpat_swap_fields");  // :538
            std::printf(" Finishing copying fields \n");
        }
        auto step2 = std::chrono::steady_clock::now();
        std::printf("Step %d takes %9.4f seconds.\n", tf, std::chrono::duration<double>(step2 - step1).count());
        step1 = step2;
        // 30 minutes before the quota: stop and dump the restart files (:558-565)
        if (std::chrono::duration<double>(step2 - wall0).count() > (cli.d("-qh") - 0.5) * 3600.0) {
            reached_quota = true;
            break;
        }
    }
    // restart files, always written when the run ends (:252-271): dump_particles(t_end), save_particle_module_state(t_end)
    // and latest_restart = the last finished frame
    if (!reached_quota) {
        tf = tf - 1;
        std::printf("Dumping restart files at the end of the simulation\n");
    } else {
        std::printf("Reached quota time. Dumping restart files.\n");
    }
    gpat_counters c{};
    gpat_get_counters(h, &c);
    {
        const std::string rdir = diag_dir + "restart/";
        ::mkdir(rdir.c_str(), 0777);
        std::vector<gpat_particle> ptl((size_t)std::max<int64_t>(c.nptl_current, 1));
        int64_t n = 0;
        CK(gpat_download_particles(h, ptl.data(), c.nptl_current, &n), "gpat_download_particles");
        char name[64];
        std::snprintf(name, sizeof(name), "particles_%04d.bin", t_end);
        FILE* f = std::fopen((rdir + name).c_str(), "wb");
        if (!f) return die(h, "write restart particles", -1);
        std::fwrite(&n, sizeof(n), 1, f);
        std::fwrite(ptl.data(), sizeof(gpat_particle), (size_t)n, f);
        std::fclose(f);
        std::snprintf(name, sizeof(name), "particle_module_state_%04d.bin", t_end);
        f = std::fopen((rdir + name).c_str(), "wb");
        if (!f) return die(h, "write restart particle_module_state", -1);
        std::fwrite(&c, sizeof(c), 1, f);
        std::fclose(f);
        const int32_t t32 = tf;
        f = std::fopen((rdir + "latest_restart").c_str(), "wb");
        if (!f) return die(h, "write restart/latest_restart", -1);
        std::fwrite(&t32, sizeof(t32), 1, f);
        std::fclose(f);
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    std::printf("Total particle steps: %llu in %.3f s (%.4g steps/s); nptl_current = %lld\n",
                (unsigned long long)total_steps, wall, total_steps / wall, (long long)c.nptl_current);
    gpat_finalize(h);
    return 0;
}
