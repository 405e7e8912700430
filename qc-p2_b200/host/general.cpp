// general.cpp -- input.json / .xyz readers (schema of the reference's src/general.cpp).
#include "general.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace {

// Flat JSON object reader: {"key": "string" | number, ...}.  That is all input.json contains
// (input.json:1-11); nested values are rejected.
std::map<std::string, std::string> parse_flat_json(const std::string &text) {
    std::map<std::string, std::string> kv;
    size_t p = 0;
    auto skip = [&] {
        while (p < text.size() && std::isspace(static_cast<unsigned char>(text[p]))) ++p;
    };
    auto fail = [&](const std::string &why) -> void { throw std::invalid_argument("input.json: " + why); };
    auto read_string = [&]() {
        std::string s;
        ++p;// opening quote
        while (p < text.size() && text[p] != '"') {
            if (text[p] == '\\' && p + 1 < text.size()) ++p;
            s.push_back(text[p++]);
        }
        if (p >= text.size()) fail("unterminated string");
        ++p;
        return s;
    };
    skip();
    if (p >= text.size() || text[p] != '{') fail("expected '{'");
    ++p;
    while (true) {
        skip();
        if (p < text.size() && text[p] == '}') break;
        if (p >= text.size() || text[p] != '"') fail("expected a key");
        const std::string key = read_string();
        skip();
        if (p >= text.size() || text[p] != ':') fail("expected ':' after key " + key);
        ++p;
        skip();
        std::string val;
        if (p < text.size() && text[p] == '"') {
            val = read_string();
        } else {
            while (p < text.size() && text[p] != ',' && text[p] != '}' && !std::isspace(static_cast<unsigned char>(text[p])))
                val.push_back(text[p++]);
            if (val.empty() || val[0] == '{' || val[0] == '[') fail("unsupported value for key " + key);
        }
        kv[key] = val;
        skip();
        if (p < text.size() && text[p] == ',') {
            ++p;
            continue;
        }
        skip();
        if (p < text.size() && text[p] == '}') break;
        fail("expected ',' or '}'");
    }
    return kv;
}

int atomic_number_of(const std::string &sym) {
    static const char *names[] = {"X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar"};
    std::string s = sym;
    if (!s.empty()) s[0] = static_cast<char>(std::toupper(static_cast<unsigned char>(s[0])));
    for (size_t k = 1; k < s.size(); ++k) s[k] = static_cast<char>(std::tolower(static_cast<unsigned char>(s[k])));
    for (int z = 1; z < 19; ++z)
        if (s == names[z]) return z;
    throw std::invalid_argument("unknown element symbol " + sym);
}

}// namespace

params read_config_json(const std::string &config_file) {
    std::cout << std::endl << "Reading configuration from " << config_file << std::endl;
    std::ifstream stream(config_file);
    if (!stream.good()) throw std::invalid_argument("cannot open " + config_file);
    std::stringstream ss;
    ss << stream.rdbuf();
    auto kv = parse_flat_json(ss.str());
    auto need = [&](const char *k) -> std::string {
        auto it = kv.find(k);
        if (it == kv.end()) throw std::invalid_argument(std::string("input.json: missing required key ") + k);
        return it->second;
    };
    params c;
    c.inputfile = need("filename");
    c.type = need("type");
    c.basis = need("basis");
    c.maxiter = std::atoi(need("maxiter").c_str());
    c.scf_conv = std::atof(need("scf_conv").c_str());
    c.ref = kv.count("ref") ? kv["ref"] : "RHF";
    c.charge = kv.count("charge") ? std::atoi(kv["charge"].c_str()) : 0;
    c.multiplicity = kv.count("multiplicity") ? std::atoi(kv["multiplicity"].c_str()) : 1;
    c.cc_conv = kv.count("cc_conv") ? std::atof(kv["cc_conv"].c_str()) : c.scf_conv;
    return c;
}

std::vector<Atom> read_geometry(const std::string &filename) {
    std::cout << std::endl << "Reading geometry from " << filename << std::endl;
    std::ifstream is(filename);
    if (!is.good()) throw std::invalid_argument("cannot open geometry file " + filename);
    if (filename.rfind(".xyz") == std::string::npos) throw std::invalid_argument("Only .xyz files are accepted as input");
    // XYZ: atom count, comment line, then "symbol x y z" in Angstrom.  Conversion constant:
    // CODATA 2018 bohr radius 0.529177210903 A (what current Libint's read_dotxyz uses).
    const double ang_to_bohr = 1.0 / 0.529177210903;
    std::string line;
    std::getline(is, line);
    const int natom = std::atoi(line.c_str());
    std::getline(is, line);// comment
    std::vector<Atom> atoms;
    for (int k = 0; k < natom; ++k) {
        if (!std::getline(is, line)) throw std::invalid_argument("xyz file ended early");
        std::istringstream ls(line);
        std::string sym;
        double x, y, z;
        if (!(ls >> sym >> x >> y >> z)) throw std::invalid_argument("malformed xyz line: " + line);
        atoms.push_back(Atom{atomic_number_of(sym), x * ang_to_bohr, y * ang_to_bohr, z * ang_to_bohr});
    }
    return atoms;
}

void print_geometry(const std::vector<Atom> &atoms) {
    std::cout << std::endl << "Geometry: " << std::endl;
    for (const auto &a : atoms) std::cout << a.atomic_number << " " << a.x << " " << a.y << " " << a.z << std::endl;
    std::cout << std::endl;
}
