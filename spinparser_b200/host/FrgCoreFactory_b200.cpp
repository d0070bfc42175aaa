// FrgCoreFactory_b200.cpp -- FrgCoreFactory::newFrgCore with the B200 flow cores registered.
//
// Replaces src/FrgCoreFactory.cpp:25-51 in a SpinParser build that links libpffrg. Identifiers,
// measurement construction and error behaviour are those of the reference factory; the only addition is the choice of
// the backend that executes computeStep()/finalizeStep():
//   * core option  backend="b200" | "cpu"  in the task file's <model ...> element, or
//   * environment  SPINPARSER_BACKEND=b200 | cpu  (the option wins; default "cpu", i.e. stock behaviour).
// With backend "b200" there is no silent fallback: if no usable GPU is present the constructor throws.
#include "lib/Exception.hpp"
#include "lib/InputParser.hpp"
#include "FrgCoreFactory.hpp"
#include "SU2/SU2MeasurementCorrelation.hpp"
#include "XYZ/XYZMeasurementCorrelation.hpp"
#include "TRI/TRIMeasurementCorrelation.hpp"
#include "B200FrgCore.hpp"
#include "B200MeasurementCorrelation.hpp"

FrgCore *FrgCoreFactory::newFrgCore(const std::string &identifier, const SpinModel &model, const std::vector<MeasurementSpecification> &measurements, const std::map<std::string, std::string> &options)
{
	if (identifier != "SU2" && identifier != "XYZ" && identifier != "TRI")
	{
		if (!measurements.empty()) throw Exception(Exception::Type::InitializationError, "Measurement [" + measurements.front().identifier + "]: Unknown model symmetry '" + identifier + "'.");
		throw Exception(Exception::Type::ArgumentError, "Spin model identifier '" + identifier + "' does not exist.");
	}

	std::string backend = "cpu";
	if (const char *env = getenv("SPINPARSER_BACKEND")) backend = env;
	auto chosen = options.find("backend");
	if (chosen != options.end()) backend = chosen->second;
	if (backend != "cpu" && backend != "b200") throw Exception(Exception::Type::InitializationError, "Unknown FRG core backend '" + backend + "'.");
	std::map<std::string, std::string> probe = options;
	const bool deviceMeasurement = backend == "b200" && b200::AdapterOptions::extract(probe).deviceMeasurement;

	std::vector<Measurement *> measurementObjects;
	for (const MeasurementSpecification &specification : measurements)
	{
		if (specification.identifier != "correlation") throw Exception(Exception::Type::InitializationError, "Measurement: Unknown measurement type '" + identifier + "'.");
		Measurement *m = nullptr;
		if (deviceMeasurement)
		{
			// the susceptibility integral runs on the GPU; same .obs output
			if (identifier == "SU2") m = new b200::B200MeasurementCorrelation<SU2FrgCore>(specification.output, specification.minCutoff, specification.maxCutoff, specification.defer);
			else if (identifier == "XYZ") m = new b200::B200MeasurementCorrelation<XYZFrgCore>(specification.output, specification.minCutoff, specification.maxCutoff, specification.defer);
			else m = new b200::B200MeasurementCorrelation<TRIFrgCore>(specification.output, specification.minCutoff, specification.maxCutoff, specification.defer);
		}
		else if (identifier == "SU2") m = new SU2MeasurementCorrelation(specification.output, specification.minCutoff, specification.maxCutoff, specification.defer);
		else if (identifier == "XYZ") m = new XYZMeasurementCorrelation(specification.output, specification.minCutoff, specification.maxCutoff, specification.defer);
		else m = new TRIMeasurementCorrelation(specification.output, specification.minCutoff, specification.maxCutoff, specification.defer);
		Log::log << Log::LogLevel::Info << "Added measurement [correlation]." << Log::endl;
		measurementObjects.push_back(m);
	}

	if (backend == "b200")
	{
		Log::log << Log::LogLevel::Info << "FRG core backend is libpffrg (B200)." << Log::endl;
		if (identifier == "SU2") return new b200::B200FrgCore<SU2FrgCore>(model, measurementObjects, options);
		if (identifier == "XYZ") return new b200::B200FrgCore<XYZFrgCore>(model, measurementObjects, options);
		return new b200::B200FrgCore<TRIFrgCore>(model, measurementObjects, options);
	}
	const std::map<std::string, std::string> stock = b200::AdapterOptions::strip(options);
	if (identifier == "SU2") return new SU2FrgCore(model, measurementObjects, stock);
	if (identifier == "XYZ") return new XYZFrgCore(model, measurementObjects, stock);
	return new TRIFrgCore(model, measurementObjects, stock);
}
